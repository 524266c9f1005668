"""GPU: the tcgen05/TMEM/TMA row-shifted GEMM against the CUDA-core backend and a torch fp32 restatement."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, W, offs, m_begin, m_end, Cin, bias, res, relu):
    a_rows = A.shape[0]
    Af, Wf = A.float(), W.float()
    out = torch.zeros(m_end - m_begin, W.shape[0], device=A.device)
    m = torch.arange(m_begin, m_end, device=A.device)
    for t, off in enumerate(offs):
        rows = m + off
        ok = (rows >= 0) & (rows < a_rows)
        a = torch.zeros(len(m), Cin, device=A.device)
        a[ok] = Af[rows[ok]]
        out += a @ Wf[:, t * Cin:(t + 1) * Cin].T
    if bias is not None:
        out += bias
    if res is not None:
        out += res.float()[m_begin:m_end]
    if relu:
        out = out.relu()
    return out


def _run(backend, A, W, offs, m_begin, m_end, Cin, Cout, bias, res, relu, fp32_out):
    from regressor_guided_image_editing_b200 import _lib
    from regressor_guided_image_editing_b200._lib import ptr, stream_ptr, check
    lib = _lib.load()
    D = torch.zeros(m_end, Cout, device=A.device, dtype=torch.float32 if fp32_out else torch.bfloat16)
    offs_c = (C.c_long * len(offs))(*offs)
    check(lib.rgie_gemm_selftest(backend, ptr(A), A.shape[0], Cin, ptr(W), W.shape[0], len(offs), offs_c, m_begin, m_end,
                                 Cout, ptr(bias), ptr(res), int(relu), ptr(D), int(fp32_out), stream_ptr(A.device)),
          "gemm_selftest")
    torch.cuda.synchronize()
    return D.float()[m_begin:m_end]


CASES = [
    # (rows, Cin, Cout, offs, m_begin, m_end_delta, bias, res, relu, fp32_out)
    (512, 64, 64, [0], 0, 0, False, False, False, True),
    (1000, 64, 64, [0], 0, 0, True, False, True, False),
    (4096, 256, 128, [0], 0, 0, True, True, True, False),
    (3000, 128, 256, [-31, -30, -29, -1, 0, 1, 29, 30, 31], 0, 0, True, False, True, False),
    (5000, 64, 16, [-(a * 227 + b) for a in range(-2, 2) for b in range(-2, 2)], 0, 0, False, False, False, True),
    (20000, 512, 512, [0], 0, 0, True, True, True, False),
    (40000, 64, 256, [0], 0, 0, True, False, False, False),
    (6000, 256, 1024, [0], 0, 0, True, True, True, False),
    (9000, 128, 128, [3000 + d for d in (-59, -58, 0, 1)], 3000, -3000, False, False, False, False),
    (2500, 512, 2048, [0], 0, 0, True, False, True, False),
    # runs of consecutive row offsets -> served from one operand slab through row-shifted matrix descriptors
    (5000, 64, 16, [a * 227 + b for a in range(-2, 2) for b in range(-1, 3)], 0, 0, False, False, False, True),
    (7000, 64, 64, [-115, -114, -113, -1, 0, 1, 113, 114, 115], 0, 0, True, False, True, False),
    (3000, 128, 128, [-3, -2, -1, 0, 1, 2, 3, 4], 0, 0, False, False, False, True),
    (4000, 256, 256, [-58 - 1, -58, -58 + 1, -1, 0, 1, 58 - 1, 58, 58 + 1], 0, 0, True, True, True, False),
]


@pytest.mark.parametrize("case", CASES, ids=[f"r{c[0]}_k{c[1]}_n{c[2]}_t{len(c[3])}" for c in CASES])
def test_tcgen05_vs_reference(case):
    rows, Cin, Cout, offs, m_begin, m_end_delta, use_bias, use_res, relu, fp32_out = case
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(rows + Cin + Cout)
    A = (torch.randn(rows, Cin, generator=g) * 0.5).to(dev).bfloat16().contiguous()
    W = (torch.randn(Cout, len(offs) * Cin, generator=g) * (1.0 / (len(offs) * Cin) ** 0.5)).to(dev).bfloat16().contiguous()
    bias = torch.randn(Cout, generator=g).to(dev) if use_bias else None
    m_end = rows + m_end_delta
    res = torch.randn(m_end, Cout, generator=g).to(dev).bfloat16().contiguous() if use_res else None
    ref = _ref(A, W, offs, m_begin, m_end, Cin, bias, res, relu)
    simt = _run(0, A, W, offs, m_begin, m_end, Cin, Cout, bias, res, relu, fp32_out)
    tc = _run(1, A, W, offs, m_begin, m_end, Cin, Cout, bias, res, relu, fp32_out)
    tol = 2e-3 if fp32_out else 2e-2      # bf16 output rounding: 2^-9 relative on values of O(1)
    err_simt = (simt - ref).abs().max().item()
    err_tc = (tc - ref).abs().max().item()
    print(f"max|simt-ref|={err_simt:.3e} max|tcgen05-ref|={err_tc:.3e} max|ref|={ref.abs().max().item():.3f}")
    assert err_simt < tol, f"CUDA-core backend off by {err_simt}"
    assert err_tc < tol, f"tcgen05 backend off by {err_tc}"
