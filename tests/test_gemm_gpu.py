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


# ---------------------------------------------------------------------------------------------------------------
# second operand (K-concatenated downsample branch), 1-bit ReLU masks, sign-bit outputs, residual ring
# ---------------------------------------------------------------------------------------------------------------
def _run_ex(backend, A, A2, a2_rows, W, offs, m_end, Cin, Cin2, Cout, bias, res, mask_bits, relu, want_bits):
    from regressor_guided_image_editing_b200 import _lib
    from regressor_guided_image_editing_b200._lib import ptr, stream_ptr, check
    lib = _lib.load()
    # guard bands around both outputs (compute-sanitizer is not available on the pool): a kernel that writes outside
    # [0, m_end) x Cout or outside the blocked bit words would disturb the sentinel
    G = 256
    Dg = torch.full((m_end + 2 * G, Cout), -7.0, device=A.device, dtype=torch.bfloat16)
    D = Dg[G:G + m_end]
    D.zero_()
    nbw = (m_end + 31) // 32 * 32 * (Cout // 32)
    Dbg = torch.full((nbw + 2 * G,), 0x5A5A5A5A, device=A.device, dtype=torch.int32) if want_bits else None
    Db = Dbg[G:G + nbw] if want_bits else None
    if want_bits:
        Db.zero_()
    offs_c = (C.c_long * len(offs))(*offs)
    check(lib.rgie_gemm_selftest_ex(backend, ptr(A), A.shape[0], Cin, ptr(A2), a2_rows, Cin2, ptr(W), W.shape[0], len(offs),
                                    offs_c, 0, m_end, Cout, ptr(bias), ptr(res), ptr(mask_bits), int(relu), ptr(D), 0,
                                    ptr(Db), stream_ptr(A.device)), "gemm_selftest_ex")
    torch.cuda.synchronize()
    assert (Dg[:G] == -7.0).all() and (Dg[G + m_end:] == -7.0).all(), f"backend {backend}: write outside the output rows"
    if want_bits:
        assert (Dbg[:G] == 0x5A5A5A5A).all() and (Dbg[G + nbw:] == 0x5A5A5A5A).all(), f"backend {backend}: write outside the bit words"
    return D.float(), Db


def _blocked_index(rows, words, dev):
    """common.cuh: bits_index(m, w, words) = ((m >> 5) * words + w) * 32 + (m & 31)."""
    m = torch.arange(rows, device=dev)[:, None]
    w = torch.arange(words, device=dev)[None, :]
    return ((m >> 5) * words + w) * 32 + (m & 31)


def _unpack_bits(flat, rows, ncols):
    words = flat.to(torch.int64)[_blocked_index(rows, ncols // 32, flat.device)] & 0xFFFFFFFF
    sh = torch.arange(32, device=flat.device)
    return ((words[:, :, None] >> sh) & 1).reshape(rows, -1).bool()


def _pack_bits(keep):
    rows, ncols = keep.shape
    sh = torch.arange(32, device=keep.device, dtype=torch.int64)
    words = (keep.reshape(rows, ncols // 32, 32).to(torch.int64) << sh).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words).to(torch.int32)
    flat = torch.zeros((rows + 31) // 32 * 32 * (ncols // 32), device=keep.device, dtype=torch.int32)
    flat[_blocked_index(rows, ncols // 32, keep.device).flatten()] = words.flatten()
    return flat.contiguous()


EX_CASES = [
    # (rows, Cin, Cin2, a2_rows, Cout, offs, res, mask, out_bits)
    (9000, 64, 64, 9000, 256, [0], False, False, True),          # layer1.0 conv3 + downsample
    (5000, 128, 256, 1250, 256, [0], False, True, False),        # stride-2 conv1 dgrad: A2 covers plane 0 only
    (4096, 512, 1024, 4000, 2048, [0], False, False, True),
    (7000, 64, 0, 0, 256, [0], True, True, True),                # residual ring + bits, single k-block tiles
    (40000, 256, 0, 0, 1024, [0], True, False, True),            # residual ring, many tiles per CTA
    (3000, 64, 0, 0, 64, [-59, -58, -57, -1, 0, 1, 57, 58, 59], False, True, True),
    (6000, 512, 0, 0, 128, [0], False, True, True),
    # CTA-pair kernel (cta_group::2; contraction >= 768, 256-wide tiles): odd tile count, residual + bits
    (9000, 1024, 0, 0, 512, [0], True, False, True),
    # CTA pairs, nine row-shifted taps, ReLU bit mask
    (5000, 128, 0, 0, 256, [-71, -70, -69, -1, 0, 1, 69, 70, 71], False, True, True),
    # CTA pairs, second operand ends inside a pair (tile 10 is below a2_rows, its partner tile 11 is not)
    (5000, 512, 512, 1400, 256, [0], False, True, False),
    # CTA pairs with 128-wide tiles (64 weight rows per CTA): the 3x3 convs of layer2, and a residual + bits case
    (8100, 128, 0, 0, 128, [-91, -90, -89, -1, 0, 1, 89, 90, 91], False, True, True),
    (5000, 1024, 0, 0, 128, [0], True, False, True),
]


@pytest.mark.parametrize("case", EX_CASES, ids=[f"r{c[0]}_k{c[1]}+{c[2]}_n{c[4]}_t{len(c[5])}" for c in EX_CASES])
def test_second_operand_and_bit_masks(case):
    rows, Cin, Cin2, a2_rows, Cout, offs, use_res, use_mask, want_bits = case
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(7 * rows + Cin + Cout)
    K = len(offs) * Cin + Cin2
    A = (torch.randn(rows, Cin, generator=g) * 0.5).to(dev).bfloat16().contiguous()
    A2 = (torch.randn(max(a2_rows, 1), max(Cin2, 1), generator=g) * 0.5).to(dev).bfloat16().contiguous() if Cin2 else None
    W = (torch.randn(Cout, K, generator=g) / K ** 0.5).to(dev).bfloat16().contiguous()
    bias = (torch.randn(Cout, generator=g) * 0.1).to(dev)
    res = torch.randn(rows, Cout, generator=g).to(dev).bfloat16().contiguous() if use_res else None
    keep = (torch.rand(rows, Cout, generator=g) > 0.4).to(dev) if use_mask else None
    mbits = _pack_bits(keep) if use_mask else None
    ref = _ref(A, W[:, :len(offs) * Cin].contiguous(), offs, 0, rows, Cin, bias, res, False)
    if Cin2:
        ref[:a2_rows] += A2.float()[:a2_rows] @ W.float()[:, len(offs) * Cin:].T
    ref = ref.relu()
    if use_mask:
        ref = ref * keep
    for backend in (0, 1):
        out, bits = _run_ex(backend, A, A2, a2_rows, W, offs, rows, Cin, Cin2, Cout, bias, res, mbits, True, want_bits)
        err = (out - ref).abs().max().item()
        assert err < 2e-2, f"backend {backend} off by {err}"
        if want_bits:
            got = _unpack_bits(bits, rows, Cout)
            assert torch.equal(got, out > 0), f"backend {backend}: sign bits differ from the stored values"


# =================================================================================================================
# fp32 parity mode on the tensor cores: gemm_tc32_kernel (exact bf16x3 operand split, six partial products) against a
# float64 restatement and against the CUDA-core fp32 kernel
# =================================================================================================================
def _ref64(A, a_rows, a_ld, Cin, A2, W, offs, m_begin, m_end, bias, res, mask, relu):
    """float64 restatement; A may have overlapping rows (row m = the Cin elements starting at element m * a_ld)."""
    dev = A.device
    flat = A.double().flatten()
    Wd = W.double().to(dev)
    m = torch.arange(m_begin, m_end, device=dev)
    out = torch.zeros(len(m), W.shape[0], dtype=torch.float64, device=dev)
    cols = torch.arange(Cin, device=dev)
    for t, off in enumerate(offs):
        rows = m + off
        ok = (rows >= 0) & (rows < a_rows)
        a = torch.zeros(len(m), Cin, dtype=torch.float64, device=dev)
        if a_ld:
            a[ok] = flat[(rows[ok] * a_ld)[:, None] + cols[None, :]]
        else:
            a[ok] = A.double()[rows[ok]]
        out += a @ Wd[:, t * Cin:(t + 1) * Cin].T
    if A2 is not None:
        ok = m < A2.shape[0]
        a = torch.zeros(len(m), A2.shape[1], dtype=torch.float64, device=dev)
        a[ok] = A2.double()[m[ok]]
        out += a @ Wd[:, len(offs) * Cin:].T
    if bias is not None:
        out += bias.double()
    if res is not None:
        out += res.double()[m_begin:m_end]
    if relu:
        out = out.relu()
    if mask is not None:
        out = out * (mask[m_begin:m_end] > 0)
    return out


def _run_fp32(backend, A, a_rows, Cin, a_ld, A2, W, offs, m_begin, m_end, Cout, bias, res, mask, relu):
    from regressor_guided_image_editing_b200 import _lib
    from regressor_guided_image_editing_b200._lib import ptr, stream_ptr, check
    lib = _lib.load()
    D = torch.zeros(m_end, Cout, device=A.device, dtype=torch.float32)
    offs_c = (C.c_long * len(offs))(*offs)
    Wh = W.float().cpu().contiguous()
    check(lib.rgie_gemm_selftest_fp32(backend, ptr(A), a_rows, Cin, a_ld, ptr(A2), 0 if A2 is None else A2.shape[0],
                                      0 if A2 is None else A2.shape[1], C.c_void_p(Wh.data_ptr()), len(offs), offs_c, m_begin,
                                      m_end, Cout, ptr(bias), ptr(res), ptr(mask), int(relu), ptr(D), stream_ptr(A.device)),
          "gemm_selftest_fp32")
    torch.cuda.synchronize()
    return D[m_begin:m_end]


FP32_CASES = [
    # (rows, Cin, Cout, offs, m_begin, m_end_delta, bias, res, relu, mask, Cin2, a_ld)
    (700, 64, 64, [0], 0, 0, False, False, False, False, 0, 0),
    (3000, 256, 128, [0], 0, 0, True, True, True, False, 0, 0),
    (2500, 128, 256, [-31, -30, -29, -1, 0, 1, 29, 30, 31], 0, 0, True, False, True, False, 0, 0),
    (5000, 64, 16, [-(a * 227 + b) for a in range(-2, 2) for b in range(-2, 2)], 0, 0, False, False, False, False, 0, 0),
    (4000, 128, 512, [0], 0, 0, True, False, True, False, 256, 0),          # second operand: K-concatenated downsample branch
    (3500, 64, 256, [0], 0, 0, False, True, False, True, 0, 0),            # input-gradient form: residual + ReLU mask
    (6000, 64, 64, [-2 * 228, -228, 0, 228], 0, 0, True, False, True, False, 0, 16),   # conv1: overlapped 16-channel rows
    (9000, 128, 128, [3000 + d for d in (-59, -58, 0, 1)], 3000, -3000, False, False, False, False, 0, 0),
    (1500, 512, 2048, [0], 0, 0, True, False, True, False, 0, 0),
]


@pytest.mark.parametrize("case", FP32_CASES, ids=[f"fp32-{i}" for i in range(len(FP32_CASES))])
def test_fp32_tensor_core_gemm_matches_float64(case):
    """|error| of the bf16x3 tensor-core GEMM against float64 must be at the level of the CUDA-core fp32 kernel's own error
    (both accumulate in fp32): <= 8x the CUDA-core error + 1e-6 of the output scale, and <= 5e-6 relative to the scale."""
    rows, Cin, Cout, offs, m_begin, m_end_delta, use_bias, use_res, relu, use_mask, Cin2, a_ld = case
    g = torch.Generator().manual_seed(rows + Cin + Cout)
    dev = "cuda"
    m_end = rows + m_end_delta
    if a_ld:
        flat = torch.randn(rows * a_ld + Cin, generator=g)
        A = flat.to(dev)
        a_rows = rows
    else:
        A = torch.randn(rows, Cin, generator=g).to(dev)
        a_rows = rows
    A2 = torch.randn(rows // 2, Cin2, generator=g).to(dev) if Cin2 else None
    W = torch.randn(Cout, len(offs) * Cin + Cin2, generator=g) / (len(offs) * Cin + Cin2) ** 0.5
    bias = torch.randn(Cout, generator=g).to(dev) if use_bias else None
    res = torch.randn(m_end, Cout, generator=g).to(dev) if use_res else None
    mask = torch.randn(m_end, Cout, generator=g).to(dev) if use_mask else None
    ref = _ref64(A, a_rows, a_ld, Cin, A2, W, offs, m_begin, m_end, bias, res, mask, relu)
    out_tc = _run_fp32(2, A, a_rows, Cin, a_ld, A2, W, offs, m_begin, m_end, Cout, bias, res, mask, relu).double()
    out_cc = _run_fp32(0, A, a_rows, Cin, a_ld, A2, W, offs, m_begin, m_end, Cout, bias, res, mask, relu).double()
    scale = ref.abs().max().item()
    e_tc, e_cc = (out_tc - ref).abs().max().item(), (out_cc - ref).abs().max().item()
    print(f"{case[:3]}: scale {scale:.3f}  tensor-core fp32 err {e_tc:.2e}  CUDA-core fp32 err {e_cc:.2e}")
    assert e_tc <= 5e-6 * scale, (e_tc, scale)
    assert e_tc <= 8 * e_cc + 1e-6 * scale
