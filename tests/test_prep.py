"""Pre-process: oracle vs torch / the reference's golden outputs (CPU), CUDA kernel vs oracle (GPU).

Tolerances.  The interpolation arithmetic lives in PyTorch and is not one function: ATen's CPU kernel rounds
differently from size to size (its compiler contracts a*b + c*d into an FMA one way, the other way or not at all
depending on which TensorIterator loop instance a shape lands in -- 37x50 -> 64x96 and 37x50 -> 48x65 differ), and
infer.py runs the CUDA kernel anyway (use_cuda=True).  The oracle therefore restates ONE fixed sequence -- the CUDA
kernel's expression with nvcc's default contraction, which is also what ATen's CPU kernel does for the 544x544 upscales
-- and is pinned against torch / the reference's golden outputs to RESIZE_ATOL (the largest effect of a 1-ulp
difference in the source coordinate on 0..255 data), far inside the north star's 1e-3 after the /255.  The CUDA kernel
is compared with the oracle BIT-EXACTLY.
"""
import numpy as np
import pytest
import torch

from tests.common import GOLDEN
from oracle.prep_oracle import fast_transform_oracle, pad_oracle, short_edge_size, bilinear_resize

RESIZE_ATOL = 4e-3          # on 0..255 pixel values; divided by std after normalisation

GOLD_CASES = {
    'resize': dict(size=(64, 96), mean=(0, 0, 0), std=(255, 255, 255)),
    'short': dict(short=(48, 80), mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375)),
    'down': dict(size=(64, 64), mean=(0, 0, 0), std=(255, 255, 255)),
}


def _oracle_case(img, spec):
    size = spec.get('size') or short_edge_size(img.shape[1], img.shape[2], *spec['short'])
    return pad_oracle(fast_transform_oracle(img, size, spec['mean'], spec['std']))


@pytest.mark.parametrize('name', sorted(GOLD_CASES))
def test_oracle_matches_reference_golden(name):
    g = np.load(GOLDEN + '/prep_small.npz')
    out, info = _oracle_case(g[name + '_in'], GOLD_CASES[name])
    assert info == g[name + '_pad'].tolist()
    assert np.abs(out - g[name + '_out']).max() <= RESIZE_ATOL / min(GOLD_CASES[name]['std'])
    if name == 'resize':                       # this shape lands on the same rounding sequence: bit-exact
        assert np.array_equal(out, g[name + '_out'])


@pytest.mark.parametrize('shape', [(480, 640, 544, 544), (1080, 1920, 544, 544), (100, 37, 544, 544), (333, 500, 608, 928),
                                   (37, 50, 48, 65), (150, 211, 64, 64), (7, 9, 5, 3)])
def test_oracle_bilinear_matches_torch(shape):
    h, w, oh, ow = shape
    rng = np.random.default_rng(h * 7 + w)
    x = rng.integers(0, 256, (1, 3, h, w)).astype(np.float32)
    ref = torch.nn.functional.interpolate(torch.from_numpy(x), size=(oh, ow), mode='bilinear', align_corners=False).numpy()
    got = bilinear_resize(x, oh, ow)
    assert np.abs(got - ref).max() <= RESIZE_ATOL
    if oh == ow == 544:                        # the north-star input size: torch CPU rounds exactly like the oracle
        assert np.array_equal(got, ref)


def test_host_mirror_geometry_and_errors():
    import orienmask_b200 as ob
    from orienmask_b200.transform import pad_geometry
    t = ob.FastCOCOTransform([dict(type='ShortEdgeResize', short_length=48, max_size=80),
                              dict(type='Normalize', mean=(0, 0, 0), std=(255, 255, 255))])
    assert t.output_size(37, 50) == short_edge_size(37, 50, 48, 80) == (48, 65)
    assert pad_geometry(48, 65) == [15, 16, 8, 8, 64, 96]                       # infer.py:22-29
    with pytest.raises(RuntimeError):
        t(torch.zeros(1, 37, 50, 3))                                            # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        ob.FastCOCOTransform([dict(type='Resize', size=8, interpolation='nearest')])
    # trainer/builder.py:108-115 builds every pipeline item as getattr(transform_class, type)(**kwargs)
    cfg = [dict(type='Resize', size=(544, 544), interpolation='bilinear', align_corners=False),
           dict(type='Normalize', mean=(0, 0, 0), std=(255, 255, 255))]                    # config/base.py:158-164
    built = ob.FastCOCOTransform(pipeline=[getattr(ob.FastCOCOTransform, c.pop('type'))(**c) for c in [dict(i) for i in cfg]], use_cuda=True)
    assert built.resize == ('fixed', (544, 544)) and built.std == (255.0, 255.0, 255.0)
    x, info = ob.pad(torch.zeros(1, 3, 48, 65))
    assert tuple(x.shape) == (1, 3, 64, 96) and info == [15, 16, 8, 8, 64, 96]


# ---- GPU ---------------------------------------------------------------------------------------
def _transform(spec):
    import orienmask_b200 as ob
    first = dict(type='Resize', size=spec['size']) if 'size' in spec else \
        dict(type='ShortEdgeResize', short_length=spec['short'][0], max_size=spec['short'][1])
    return ob.FastCOCOTransform([first, dict(type='Normalize', mean=spec['mean'], std=spec['std'])])


@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(GOLD_CASES))
@pytest.mark.parametrize('dtype', [torch.uint8, torch.float32])
def test_kernel_matches_reference_golden(name, dtype):
    import orienmask_b200 as ob
    g = np.load(GOLDEN + '/prep_small.npz')
    img = torch.from_numpy(g[name + '_in']).to('cuda:0', dtype)
    tr = _transform(GOLD_CASES[name])
    out, info = tr.transform_and_pad(img)
    assert info == g[name + '_pad'].tolist()
    assert np.abs(out.cpu().numpy() - g[name + '_out']).max() <= RESIZE_ATOL / min(GOLD_CASES[name]['std'])
    ref, _ = _oracle_case(g[name + '_in'], GOLD_CASES[name])
    assert np.array_equal(out.cpu().numpy(), ref)                               # kernel == oracle, bit for bit
    two_step, info2 = ob.pad(tr(img))                                           # infer.py:149-150 call shape
    assert info2 == info and torch.equal(two_step, out)


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(480, 640, 544, 544), (427, 640, 544, 544), (720, 1280, 960, 960), (31, 45, 33, 47)])
def test_kernel_matches_oracle_full_size(shape):
    h, w, oh, ow = shape
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    spec = dict(size=(oh, ow), mean=(0, 0, 0), std=(255, 255, 255))
    ref, info = pad_oracle(fast_transform_oracle(img, spec['size'], spec['mean'], spec['std']))
    out, got_info = _transform(spec).transform_and_pad(torch.from_numpy(img).cuda())
    assert got_info == info
    assert np.array_equal(out.cpu().numpy(), ref)
    # strided batch view (images taken from a larger buffer)
    big = torch.from_numpy(np.concatenate([img, img], 0)).cuda()[::2]
    out2, _ = _transform(spec).transform_and_pad(big)
    assert np.array_equal(out2.cpu().numpy()[0], ref[0])


@pytest.mark.gpu
def test_identity_resize_is_exact_division():
    """544x544 sources (the bench workload): the resize is the identity and the output is exactly v / 255."""
    img = torch.randint(0, 256, (2, 544, 544, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(2))
    out = _transform(dict(size=(544, 544), mean=(0, 0, 0), std=(255, 255, 255)))(img.cuda())
    assert torch.equal(out.cpu(), img.permute(0, 3, 1, 2).float() / 255.0)
