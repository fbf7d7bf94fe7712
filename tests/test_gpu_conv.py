"""GPU numerics of the convolution engines (C-ABI om_conv_*) against a plain PyTorch fp64 reference."""
import pytest
import torch

from tests.common import run_engine_conv, torch_conv_ref

pytestmark = pytest.mark.gpu

F32, F16, SPLIT = 0, 1, 2
ACT, PARTIAL, NCHW = 0, 1, 2

# (B, cin, cout, H, W, k, stride, kind, residual, upadd)
CASES = [
    (2, 64, 128, 16, 16, 1, 1, ACT, False, False),
    (2, 64, 128, 17, 34, 3, 1, ACT, False, False),
    (1, 128, 256, 34, 34, 3, 1, ACT, True, False),
    (3, 32, 64, 32, 32, 3, 2, ACT, False, False),
    (2, 64, 32, 16, 16, 1, 1, ACT, False, False),
    (2, 32, 64, 16, 16, 3, 1, ACT, True, False),
    (2, 128, 256, 20, 12, 3, 2, ACT, False, False),
    (2, 512, 1024, 6, 6, 3, 1, ACT, False, False),
    (2, 256, 255, 8, 8, 1, 1, NCHW, False, False),
    (1, 256, 18, 24, 24, 1, 1, NCHW, False, False),
    (2, 64, 128, 12, 12, 1, 1, PARTIAL, False, True),
    (2, 512, 256, 12, 12, 1, 1, ACT, False, True),
    (1, 128, 256, 136, 136, 3, 1, ACT, False, False),
    (8, 128, 256, 136, 136, 3, 1, ACT, False, False),     # eight tiles per CTA pair: halo stages (split precision: lo / hi half stages) recycle
    (1, 128, 256, 68, 68, 3, 1, ACT, True, False),        # halo tiles with a partial last tile column (68 = 8*8 + 4)
    (8, 128, 256, 68, 68, 3, 1, ACT, True, False),        # ... several tiles per CTA, N = 256: 64-column residual chunks in two shared buffers
    (2, 64, 64, 24, 72, 3, 1, ACT, False, False),
    (3, 32, 64, 40, 16, 3, 1, ACT, True, False),          # 64-byte pixels (SWIZZLE_64B halo)
    (2, 256, 128, 17, 17, 1, 1, ACT, False, False),
    # flat (im2col-gathered) tiles: runs of 128 pixels that cross rows and images, odd tile counts, strided traversal
    (5, 64, 64, 17, 17, 3, 1, ACT, True, False),
    (7, 128, 64, 17, 17, 1, 1, ACT, False, False),
    (4, 64, 128, 34, 34, 3, 2, ACT, False, False),
    (3, 256, 512, 34, 34, 3, 1, ACT, True, False),
    # ... enough pixel tiles for N = 256 and several waves: with and without a residual, stride 2
    (16, 256, 512, 34, 34, 3, 1, ACT, True, False),
    (32, 512, 1024, 17, 17, 3, 1, ACT, False, False),
    (16, 256, 512, 68, 68, 3, 2, ACT, False, False),
    (3, 96, 64, 10, 14, 3, 1, ACT, False, False),         # 64-byte pixels (BK = 32), three chunks
    (2, 64, 128, 34, 34, 1, 1, PARTIAL, False, False),
]


def _data(case, seed=0):
    B, cin, cout, H, W, k, stride, kind, use_res, use_up = case
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, cin, H, W, generator=g).cuda()
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).cuda()
    b = torch.randn(cout, generator=g).cuda() if kind != PARTIAL else None
    Ho, Wo = H // stride, W // stride
    res = torch.randn(B, cout, Ho, Wo, generator=g).cuda() if use_res else None
    up = torch.randn(B, cout, Ho // 2, Wo // 2, generator=g).cuda() if use_up else None
    return x, w, b, res, up


@pytest.mark.parametrize('case', CASES)
def test_f32_engine(case):
    B, cin, cout, H, W, k, stride, kind, _, _ = case
    x, w, b, res, up = _data(case)
    leaky = kind == ACT
    got = run_engine_conv(x, w, b, stride, leaky, kind, res, up, precision=F32)
    ref = torch_conv_ref(x, w, b, stride, leaky, kind, res, up)
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4), float((got - ref).abs().max())


@pytest.mark.parametrize('case', CASES)
def test_tcgen05_engine(case):
    B, cin, cout, H, W, k, stride, kind, _, _ = case
    x, w, b, res, up = _data(case)
    leaky = kind == ACT
    got = run_engine_conv(x, w, b, stride, leaky, kind, res, up, precision=F16)
    # operands rounded to fp16 exactly as the engine stores them; accumulation is fp32 in both
    ref = torch_conv_ref(x, w, b, stride, leaky, kind, res, up, quantize=True)
    tol = 2e-3 if kind != ACT else 1e-2           # fp16 output rounding for activation outputs
    err = float((got - ref).abs().max())
    assert torch.allclose(got, ref, atol=tol, rtol=4e-3), err


def _split_check(got, ref):
    """fp32-grade: every product is carried by fp16 hi + lo pairs (~22 bits) and accumulated in fp32 on the tensor cores."""
    err = float((got - ref).abs().max())
    rel = float((got - ref).norm() / ref.norm())
    assert torch.allclose(got, ref, atol=2e-4, rtol=1e-4) and rel < 1e-5, (err, rel)


@pytest.mark.parametrize('case', CASES)
def test_split_precision_engine(case):
    """OM_PREC_SPLIT (the tensor-core parity mode) against fp64 torch on the UNROUNDED fp32 operands, at the fp32 engine's tolerance."""
    B, cin, cout, H, W, k, stride, kind, _, _ = case
    x, w, b, res, up = _data(case)
    leaky = kind == ACT
    got = run_engine_conv(x, w, b, stride, leaky, kind, res, up, precision=SPLIT)
    _split_check(got, torch_conv_ref(x, w, b, stride, leaky, kind, res, up))


def test_split_precision_weight_scale_and_small_values():
    """Weights far below fp16's normal range (1e-6) and activations spanning 1e-3 .. 1e3: the power-of-two weight scale keeps W_lo
    normal, and the hi + lo activation pair keeps small values next to large ones."""
    g = torch.Generator().manual_seed(21)
    x = (torch.randn(2, 64, 17, 17, generator=g) * torch.logspace(-3, 3, 64).view(1, 64, 1, 1)).cuda()
    w = (torch.randn(128, 64, 3, 3, generator=g) * 1e-6).cuda()
    b = (torch.randn(128, generator=g) * 1e-4).cuda()
    got = run_engine_conv(x, w, b, 1, True, ACT, precision=SPLIT)
    ref = torch_conv_ref(x, w, b, 1, True, ACT)
    assert float((got - ref).norm() / ref.norm()) < 1e-5


@pytest.mark.parametrize('case', [c for c in CASES if c[9]] + [(2, 128, 128, 24, 16, 1, 1, ACT, False, True), (1, 64, 256, 14, 34, 1, 1, PARTIAL, False, True)])
def test_split_precision_upadd_staged_by_tma(case):
    B, cin, cout, H, W, k, stride, kind, _, _ = case
    x, w, b, res, up = _data(case, seed=3)
    leaky = kind == ACT
    got = run_engine_conv(x, w, b, stride, leaky, kind, res, up, precision=SPLIT, extra_rows=2)
    _split_check(got, torch_conv_ref(x, w, b, stride, leaky, kind, res, up))


@pytest.mark.parametrize('case', [c for c in CASES if c[9]] + [(2, 128, 128, 24, 16, 1, 1, ACT, False, True), (1, 64, 256, 14, 34, 1, 1, PARTIAL, False, True)])
def test_tcgen05_engine_upadd_staged_by_tma(case):
    """rows_per_image(out) == 2 * rows_per_image(partial): the up-add source tile is staged by TMA (model geometry)."""
    B, cin, cout, H, W, k, stride, kind, _, _ = case
    x, w, b, res, up = _data(case, seed=3)
    leaky = kind == ACT
    got = run_engine_conv(x, w, b, stride, leaky, kind, res, up, precision=F16, extra_rows=2)
    ref = torch_conv_ref(x, w, b, stride, leaky, kind, res, up, quantize=True)
    tol = 2e-3 if kind != ACT else 1e-2
    assert torch.allclose(got, ref, atol=tol, rtol=4e-3), float((got - ref).abs().max())


@pytest.mark.parametrize('cout', [75, 18, 255])
def test_head_channel_counts_pad_to_32(cout):
    """NCHW heads of any width: 20 classes -> 3 * 25 = 75 channels -> an N = 96 pair tile (48 weight rows per CTA)."""
    g = torch.Generator().manual_seed(cout)
    x = torch.randn(2, 256, 12, 20, generator=g).cuda()
    w = (torch.randn(cout, 256, 1, 1, generator=g) / 16).cuda()
    b = torch.randn(cout, generator=g).cuda()
    for prec, quant in ((F16, True), (SPLIT, False)):
        got = run_engine_conv(x, w, b, 1, False, NCHW, precision=prec)
        ref = torch_conv_ref(x, w, b, 1, False, NCHW, quantize=quant)
        assert torch.allclose(got, ref, atol=2e-3 if quant else 2e-4, rtol=4e-3 if quant else 1e-4), (prec, float((got - ref).abs().max()))


@pytest.mark.parametrize('precision', [F32, F16, SPLIT])
def test_parity_split_layouts(precision):
    """in_s2d (stride-2 layer reading a parity-split input) and out_s2d (layer writing one), both engines."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 32, 32, 48, generator=g).cuda()
    w = (torch.randn(64, 32, 3, 3, generator=g) / (32 * 9) ** 0.5).cuda()
    b = torch.randn(64, generator=g).cuda()
    q = precision == F16
    tol = dict(atol=1e-2, rtol=4e-3) if q else dict(atol=2e-4, rtol=1e-4)
    # extra_rows=2 keeps rows_per_image even for both the stride-2 input (2x) and the s2d output
    got = run_engine_conv(x, w, b, 2, True, ACT, precision=precision, extra_rows=2, in_s2d=True)
    assert torch.allclose(got, torch_conv_ref(x, w, b, 2, True, ACT, quantize=q), **tol)
    res = torch.randn(2, 64, 32, 48, generator=g).cuda()
    got = run_engine_conv(x, w, b, 1, True, ACT, res, precision=precision, extra_rows=2, out_s2d=True)
    assert torch.allclose(got, torch_conv_ref(x, w, b, 1, True, ACT, res, quantize=q), **tol)
