"""CPU check of the scale filter parameter-gradient expression of csrc/filters.cu (scale_kernel<true>) against autograd
through the oracle, in fp32 and fp64:  python tools/scale_kink_check.py sx sy cx cy
At generic values the three agree to ~3e-5; at (1.05, 1.03, 3, 5) every 21st column samples EXACTLY on a pixel centre and the
fp32 / fp64 evaluations of the same expression differ by 4-10 percent (the bilinear kink)."""
import sys
PARAMS = tuple(float(v) for v in sys.argv[1:5])
import torch, sys, math
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as O
torch.manual_seed(0)
H=W=256
im=O.synthetic_image(0,H,W)[None]
p=torch.tensor([list(PARAMS)],requires_grad=True)
imr=im.clone().requires_grad_(True)
out=O.apply_one('scale',imr,p)
out=torch.clamp(out,0,1)
g=torch.randn(out.shape,generator=torch.Generator().manual_seed(1))
gp,gi=torch.autograd.grad((out*g).sum(),[p,imr])
print('autograd d(param):',gp)
# kernel formula in float64 and float32
def kernel_formula(dtype):
    sx,sy,cx,cy=[torch.tensor(v,dtype=dtype) for v in PARAMS]
    a=2.0/(W-1); b=2.0/(H-1)
    tx=(1-sx)*cx; ty=(1-sx)*cy
    inv_sx=1/sx; inv_sy=1/sy
    t02=-(sx+a*tx-1)/sx; t12=-(sy+b*ty-1)/sy
    xn=torch.linspace(-1,1,W,dtype=dtype); yn=torch.linspace(-1,1,H,dtype=dtype)
    gx=xn*inv_sx+t02; gy=yn*inv_sy+t12
    ix=((gx+1)*0.5)*(W-1); iy=((gy+1)*0.5)*(H-1)
    x0=torch.floor(ix).long(); y0=torch.floor(iy).long()
    wx1=(ix-x0).to(dtype); wy1=(iy-y0).to(dtype)
    imd=im[0].to(dtype)
    def at(yy,xx):
        vy=(yy>=0)&(yy<H); vx=(xx>=0)&(xx<W)
        yyc=yy.clamp(0,H-1); xxc=xx.clamp(0,W-1)
        v=imd[:,yyc][:,:,xxc]
        return v*(vy[:,None]&vx[None,:]).to(dtype)
    v00=at(y0,x0); v01=at(y0,x0+1); v10=at(y0+1,x0); v11=at(y0+1,x0+1)
    wx0=1-wx1; wy0=1-wy1
    o=v00*(wy0[:,None]*wx0[None,:])+v01*(wy0[:,None]*wx1[None,:])+v10*(wy1[:,None]*wx0[None,:])+v11*(wy1[:,None]*wx1[None,:])
    gm=g[0].to(dtype)*((o>=0)&(o<=1)).to(dtype)
    gix=(gm*((v01-v00)*wy0[:,None]+(v11-v10)*wy1[:,None])).sum(0)
    giy=(gm*((v10-v00)*wx0[None,:]+(v11-v01)*wx1[None,:])).sum(0)
    ggx=gix*0.5*(W-1); ggy=giy*0.5*(H-1)
    XN=xn[None,:].expand(H,W); YN=yn[:,None].expand(H,W)
    d_sx=(ggx*((-XN+a*cx-1)/(sx*sx))+ggy*(b*cy/sy)).sum()
    d_sy=(ggy*((-YN+b*(1-sx)*cy-1)/(sy*sy))).sum()
    d_cx=(ggx*(-a*(1-sx)/sx)).sum()
    d_cy=(ggy*(-b*(1-sx)/sy)).sum()
    return torch.stack([d_sx,d_sy,d_cx,d_cy]), o
for dt in (torch.float64, torch.float32):
    k,o=kernel_formula(dt)
    print(dt,'kernel formula:',k.tolist(),' fwd max diff', (o.float()-O.apply_one('scale',im,p.detach())[0]).abs().max().item())
