"""CPU checks of the transform host algebra: the closed-form index->coordinate matrix the CUDA resampler consumes must
equal the dense coordinates the reference generates (F.affine_grid + grid_sample un-normalisation / create_grid @ affine)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from monai_b200.transforms import utils as U
from oracle import transforms as otr


def _dense_coords_from_affine_grid(xform, src_shape, dst_shape, align):
    r = len(src_shape)
    theta = torch.as_tensor(xform, dtype=torch.float64)[None]
    theta = otr._normalize_transform(src_shape, False) @ theta @ torch.linalg.inv(otr._normalize_transform(dst_shape, False))
    rev = list(range(r - 1, -1, -1))
    t2 = theta.clone(); t2[:, :r] = theta[:, rev]
    t3 = t2.clone(); t3[:, :, :r] = t2[:, :, rev]
    grid = F.affine_grid(t3[:, :r], [1, 1, *dst_shape], align_corners=align)[0]  # (..., xyz) normalised
    g = grid.flip(-1)
    s = torch.tensor(src_shape, dtype=torch.float64)
    return ((g + 1) / 2 * (s - 1)) if align else (((g + 1) * s - 1) / 2)


@pytest.mark.parametrize("align", [False, True])
def test_sample_matrix_from_xform_equals_affine_grid(align):
    rng = np.random.default_rng(0)
    xform = np.eye(4)
    xform[:3, :3] += rng.normal(0, 0.2, (3, 3))
    xform[:3, 3] = rng.normal(0, 3, 3)
    src, dst = (9, 11, 13), (7, 12, 10)
    want = _dense_coords_from_affine_grid(xform, src, dst, align).numpy()
    m = U.sample_matrix_from_xform(xform, src, dst, align)
    idx = np.stack(np.meshgrid(*[np.arange(d) for d in dst], indexing="ij"), -1)
    got = idx @ m[:3, :3].T + m[:3, 3]
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-9)


def test_host_helpers_match_oracle():
    rng = np.random.default_rng(1)
    a = np.eye(4)
    a[:3, :3] = np.diag([1.3, 0.7, 2.1]) @ np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    a[:3, 3] = [4, -2, 7]
    np.testing.assert_allclose(U.affine_to_spacing(a), otr.affine_to_spacing(a))
    for diag in (True, False):
        np.testing.assert_allclose(U.zoom_affine(a, (1.0, 1.0, 1.5), diag), otr.zoom_affine(a, (1.0, 1.0, 1.5), diag))
    na = U.zoom_affine(a, (1.0, 1.0, 1.5), False)
    s1, o1 = U.compute_shape_offset((20, 24, 18), a, na)
    s2, o2 = otr.compute_shape_offset((20, 24, 18), a, na)
    assert tuple(s1) == tuple(s2)
    np.testing.assert_allclose(o1, o2)
    rot = U.create_rotate(3, (0.3, -0.2, 0.1))
    np.testing.assert_array_equal(rot, otr._create_rotate3((0.3, -0.2, 0.1)).numpy())


def test_rand_affine_draw_order():
    from monai_b200.transforms import RandAffined

    t = RandAffined(keys=["image"], prob=1.0, rotate_range=(0.2,) * 3, scale_range=(0.1,) * 3, translate_range=(5,) * 3)
    t.set_random_state(seed=0)
    t.randomize(None)
    t.rand_affine.randomize()
    t.rand_affine.rand_affine_grid.matrix(3, randomize=True)
    g = t.rand_affine.rand_affine_grid
    rot, shear, trans, scale = otr.rand_affine_params(0, (0.2,) * 3, (), (5,) * 3, (0.1,) * 3)
    assert g.rotate_params == rot and g.translate_params == trans and g.scale_params == scale


def test_spacing_host_algebra_reproduces_the_reference_unit_test_shapes(golden_dir):
    """Spacing's host side (affine_to_spacing / zoom_affine / compute_shape_offset, float64 numpy) against the output shapes of
    every golden case of the reference's tests/transforms/test_spacing.py (ref_unit_goldens.npz), negative pixdims and a
    4-D spatial input included.  The resampler is stubbed: shapes are decided before any device work."""
    import json
    import os

    import numpy as np
    import torch

    from monai_b200.data import MetaTensor
    from monai_b200.transforms import Spacing

    g = np.load(os.path.join(golden_dir, "ref_unit_goldens.npz"))
    n = 0
    for rec in json.loads(str(g["index"])):
        if rec["kind"] != "spacing":
            continue
        tag = rec["tag"]
        sp = Spacing(**rec["init"])
        seen = {}

        def stub(img, dst_affine=None, spatial_size=None, **kwargs):
            seen["size"] = tuple(int(v) for v in spatial_size)
            return img

        sp.sp_resample = stub
        x = torch.from_numpy(g[tag + ".x"]).float()
        sp(MetaTensor(x, affine=torch.as_tensor(g[tag + ".affine"])), **rec["call"])
        want = g[tag + ".y"].shape
        assert seen["size"] == tuple(want[1 : 1 + len(seen["size"])]), (tag, rec["init"], seen["size"], want)
        n += 1
    assert n == 17


def test_spacingd_shapes_and_affines_of_the_reference_unit_tests(monkeypatch):
    """tests/transforms/test_spacingd.py:28-99 (six dictionary cases: 3-D, 2-D with a 3x3 affine, no metadata, per-key modes,
    keys with different affines): output shape and the MetaTensor affine after the transform.  Only the resampling kernel is
    replaced (by a zero tensor of the requested shape); SpatialResample / Spacing / Spacingd bookkeeping runs as shipped."""
    import numpy as np
    import torch

    import monai_b200.transforms.spatial as S
    from monai_b200.data import MetaTensor
    from monai_b200.transforms import Spacingd

    monkeypatch.setattr(S, "_resample", lambda img, mat, r, out_shape, mode, padding_mode, align: torch.zeros((img.shape[0], *out_shape)))
    ones = lambda *s: torch.ones(s)  # noqa: E731
    cases = [
        ({"image": MetaTensor(ones(2, 10, 15, 20), affine=torch.eye(4))}, dict(keys="image", pixdim=(1, 2, 1.4)), (2, 10, 8, 15), np.diag([1, 2, 1.4, 1.0])),
        ({"image": MetaTensor(ones(2, 10, 20), affine=torch.eye(3))}, dict(keys="image", pixdim=(1, 2)), (2, 10, 10), np.diag((1, 2, 1))),
        ({"image": MetaTensor(ones(2, 10, 20))}, dict(keys="image", pixdim=(1, 2)), (2, 10, 10), np.diag((1, 2, 1, 1))),
        ({"image": MetaTensor(torch.arange(20.0).reshape(2, 1, 10), affine=torch.eye(4)), "seg": MetaTensor(ones(2, 1, 10), affine=torch.eye(4))},
         dict(keys=("image", "seg"), mode="nearest", pixdim=(1, 0.2)), (2, 1, 46), np.diag((1, 0.2, 1, 1))),
        ({"image": MetaTensor(ones(2, 1, 10), affine=torch.eye(4)), "seg": MetaTensor(ones(2, 1, 10), affine=torch.eye(4))},
         dict(keys=("image", "seg"), mode=("bilinear", "nearest"), pixdim=(1, 0.2)), (2, 1, 46), np.diag((1, 0.2, 1, 1))),
        ({"image": MetaTensor(ones(2, 1, 10), affine=torch.eye(4)), "seg1": MetaTensor(ones(2, 1, 10), affine=torch.diag(torch.tensor([2.0, 2, 2, 1]))),
          "seg2": MetaTensor(ones(2, 1, 10), affine=torch.eye(4))},
         dict(keys=("image", "seg1", "seg2"), mode=("bilinear", "nearest", "nearest"), pixdim=(1, 1, 1)), (2, 1, 10), np.diag((1, 1, 1, 1))),
    ]
    for data, kw, shape, affine in cases:
        res = Spacingd(**kw)(data)
        for key in data:
            # a key whose SPACING differs from the first key's is resampled on its own grid (real reference: seg1 -> (2, 1, 19))
            want = (2, 1, 19) if key == "seg1" else shape
            assert tuple(res[key].shape) == want, (kw, key, tuple(res[key].shape))
        np.testing.assert_allclose(res["image"].affine.numpy(), affine, atol=1e-9, err_msg=str(kw))


def test_lazy_compose_runs_one_resample_and_matches_the_reference_affine(monkeypatch, golden_dir):
    """Compose(lazy=True): Spacingd o RandAffined record their matrices and ONE resample runs with the composed matrix; shape,
    affine and the number of applied operations equal the real reference's (fixture: tests/golden/lazy_inverse.npz)."""
    import os

    import numpy as np
    import torch

    import monai_b200.transforms.spatial as S
    from monai_b200.data import MetaTensor
    from monai_b200.transforms import Compose, RandAffined, Spacingd

    g = np.load(os.path.join(golden_dir, "lazy_inverse.npz"))
    calls = []

    def stub(img, mat, r, out_shape, mode, padding_mode, align):
        calls.append((np.asarray(mat).copy(), tuple(out_shape), mode, padding_mode))
        return torch.zeros((img.shape[0], *out_shape))

    monkeypatch.setattr(S, "_resample", stub)

    def pipe(lazy):
        c = Compose([Spacingd(keys=["image"], pixdim=(1.0, 1.0, 1.0), mode="bilinear"),
                     RandAffined(keys=["image"], prob=1.0, rotate_range=(0.2,) * 3, scale_range=(0.1,) * 3, translate_range=(5,) * 3, mode="bilinear", padding_mode="border")],
                    lazy=lazy)
        c.transforms[1].set_random_state(seed=0)
        return c

    x = MetaTensor(torch.from_numpy(g["x"]), affine=torch.as_tensor(g["x_affine"]))
    y = pipe(True)({"image": x})["image"]
    assert len(calls) == 1, "lazy mode must resample once"
    assert tuple(y.shape) == tuple(g["lazy.y"].shape) and calls[0][1] == tuple(g["lazy.y"].shape[1:])
    assert calls[0][2] == "bilinear" and calls[0][3] == "border"
    np.testing.assert_allclose(y.affine.numpy(), g["lazy.affine"], atol=1e-6)
    assert len(y.applied_operations) == int(g["lazy.n_applied"]) and not y.pending_operations
    calls.clear()
    y2 = pipe(False)({"image": MetaTensor(torch.from_numpy(g["x"]), affine=torch.as_tensor(g["x_affine"]))})["image"]
    assert len(calls) == 2
    np.testing.assert_allclose(y2.affine.numpy(), g["eager.affine"], atol=1e-6)


# ---- AffineTransform (monai/networks/layers/spatial_transforms.py:439-592): reference unit goldens, tests/networks/layers/test_affine_transform.py
_T3 = np.pi / 3
_ROT2 = [[np.cos(_T3), -np.sin(_T3), 0], [np.sin(_T3), np.cos(_T3), 0], [0, 0, 1]]
_ROT3 = [[1, 0, 0, 0], [0.0, np.cos(_T3), -np.sin(_T3), 0], [0, np.sin(_T3), np.cos(_T3), 0], [0, 0, 0, 1]]
_IMG34 = [[[[4.0, 1.0, 3.0, 2.0], [7.0, 6.0, 8.0, 5.0], [3.0, 5.0, 3.0, 6.0]]]]
AFFINE_TRANSFORM_GOLDENS = [   # (init kwargs, image, theta, call spatial_size, expected, atol)   lines 131-262 of the reference test
    (dict(align_corners=False), _IMG34, [[1.0, 0.0, 0.0], [0.0, 1.0, -1.0]], None, [[[[0, 4, 1, 3], [0, 7, 6, 8], [0, 3, 5, 3]]]], 1e-5),
    (dict(align_corners=False), _IMG34, [[1.0, 0.0, -1.0], [0.0, 1.0, -1.0]], None, [[[[0, 0, 0, 0], [0, 4, 1, 3], [0, 7, 6, 8]]]], 1e-5),
    (dict(align_corners=False), _IMG34, [[1.0, 0.0, -1.0], [0.0, 1.0, 0.0]], None, [[[[0, 0, 0, 0], [4, 1, 3, 2], [7, 6, 8, 5]]]], 1e-5),
    (dict(spatial_size=(3, 2), align_corners=False), np.arange(1.0, 13.0).reshape(1, 1, 3, 4), [[1.0, 0.0, 0.0], [0.0, 2.0, 0.0]], None,
     [[[[1, 3], [5, 7], [9, 11]]]], 1e-5),
    (dict(), np.arange(1.0, 13.0).reshape(1, 1, 3, 4), [[2.0, 0.0, 0.0], [0.0, 1.0, 0.0]], (1, 4), [[[[2.333333, 3.333333, 4.333333, 5.333333]]]], 1e-4),
    (dict(spatial_size=(1, 2)), np.arange(1.0, 13.0).reshape(1, 1, 3, 4), [[2.0, 0.0, 0.0], [0.0, 2.0, 0.0]], None, [[[[1.458333, 4.958333]]]], 1e-5),
    (dict(spatial_size=(1, 2), zero_centered=True), np.arange(1.0, 13.0).reshape(1, 1, 3, 4), [[2.0, 0.0, 0.0], [0.0, 2.0, 0.0]], None, [[[[5.5, 7.5]]]], 1e-5),
    (dict(align_corners=False), np.arange(24.0).reshape(1, 1, 4, 6), _ROT2, None,
     [[[[0.0, 0.06698727, 0.0, 0.0, 0.0, 0.0], [3.8660254, 0.86602557, 0.0, 0.0, 0.0, 0.0], [7.732051, 3.035899, 0.73205125, 0.0, 0.0, 0.0],
        [11.598076, 6.901923, 2.7631402, 0.0, 0.0, 0.0]]]], 1e-3),
    (dict(spatial_size=(3, 4), padding_mode="border", align_corners=False, mode="bilinear"), np.arange(24.0).reshape(1, 1, 4, 6), _ROT2, None,
     [[[[7.1525574e-07, 4.9999994e-01, 1.0, 1.4999999], [3.8660259, 1.3660253, 1.8660252, 2.3660252], [7.7320518, 3.0358994, 2.7320509, 3.2320507]]]], 1e-3),
    (dict(spatial_size=(3, 4, 2), padding_mode="border", align_corners=False, mode="bilinear"), np.arange(48.0).reshape(2, 1, 4, 2, 3), _ROT3, None,
     [[[[[0.00000006, 0.5000001], [2.3660254, 1.3660254], [4.732051, 2.4019241], [5.0, 3.9019237]],
        [[6.0, 6.5], [8.366026, 7.3660254], [10.732051, 8.401924], [11.0, 9.901924]],
        [[12.0, 12.5], [14.366026, 13.366025], [16.732052, 14.401924], [17.0, 15.901923]]]],
      [[[[24.0, 24.5], [26.366024, 25.366024], [28.732052, 26.401924], [29.0, 27.901924]],
        [[30.0, 30.5], [32.366028, 31.366026], [34.732048, 32.401924], [35.0, 33.901924]],
        [[36.0, 36.5], [38.366024, 37.366024], [40.73205, 38.401924], [41.0, 39.901924]]]]], 1e-4),
]


def _sample_with_index_matrix(img, m, out_shape, mode, padding_mode, align):
    """CPU stand-in for b200_resample_affine (test helper): evaluates an output-index -> source-index matrix with F.grid_sample."""
    r = len(out_shape)
    idx = np.stack(np.meshgrid(*[np.arange(d, dtype=np.float64) for d in out_shape], indexing="ij"), -1)
    coords = idx @ m[:r, :r].T + m[:r, r]
    s = np.asarray(img.shape[1:], dtype=np.float64)
    g = (coords / np.where(s > 1, s - 1, 1.0) * 2 - 1) if align else ((coords * 2 + 1) / s - 1)
    grid = torch.from_numpy(g[..., ::-1].copy())[None]
    return F.grid_sample(img[None].double(), grid, mode=mode, padding_mode=padding_mode, align_corners=align)[0]


@pytest.mark.parametrize("case", range(len(AFFINE_TRANSFORM_GOLDENS)))
def test_affine_transform_reference_unit_goldens(case):
    """The oracle restatement AND the host matrix of monai_b200.AffineTransform reproduce every golden of the reference's unit test."""
    init, image, theta, call_size, expected, atol = AFFINE_TRANSFORM_GOLDENS[case]
    image = torch.as_tensor(np.asarray(image), dtype=torch.float32)
    theta_t = torch.as_tensor(np.asarray(theta), dtype=torch.float32)
    kw = dict(normalized=False, mode="bilinear", padding_mode="zeros", align_corners=True, reverse_indexing=True, zero_centered=False)
    kw.update({k: v for k, v in init.items() if k != "spatial_size"})
    size = call_size if call_size is not None else init.get("spatial_size")
    got = otr.affine_transform(image, theta_t, spatial_size=size, **kw).numpy()
    np.testing.assert_allclose(got, np.asarray(expected), atol=atol, rtol=1e-4)
    # host matrix: one per batch item, evaluated by the CPU stand-in of the resampling kernel
    sr = image.dim() - 2
    dst = tuple(size) if size is not None else tuple(image.shape[2:])
    th = np.asarray(theta, dtype=np.float64)
    if th.shape[0] == sr:
        th = np.vstack([th, np.eye(sr + 1)[-1]])
    m = U.affine_transform_matrix(th, tuple(image.shape[2:]), dst, kw["normalized"], kw["reverse_indexing"], kw["align_corners"], kw["zero_centered"])
    for b in range(image.shape[0]):
        out = _sample_with_index_matrix(image[b], m, dst, kw["mode"], kw["padding_mode"], kw["align_corners"]).numpy()
        np.testing.assert_allclose(out, np.asarray(expected)[b], atol=atol, rtol=1e-4)


@pytest.mark.parametrize("normalized,reverse,align,zero", [(False, True, True, False), (False, True, False, True), (False, False, False, False),
                                                           (True, False, False, False), (True, True, True, False), (False, False, True, True)])
@pytest.mark.parametrize("sr", [2, 3])
def test_affine_transform_matrix_equals_the_dense_grid_of_the_reference_chain(normalized, reverse, align, zero, sr):
    """All convention switches, random theta, output size != input size: the matrix equals affine_grid + grid_sample of the restated
    reference chain on a random image (bilinear, border) to float32 round-off; test_forward_2d / 3d of the reference (normalized=True,
    reverse_indexing=False == plain F.affine_grid + F.grid_sample) are the (True, False, False, False) rows."""
    rng = np.random.default_rng(sr * 10 + normalized + 2 * reverse)
    src = (5, 7, 6)[:sr]
    dst = (6, 4, 8)[:sr]
    th = np.eye(sr + 1)
    th[:sr, :sr] += rng.normal(0, 0.25, (sr, sr))
    th[:sr, sr] = rng.normal(0, 0.3 if normalized else 1.0, sr)
    img = torch.from_numpy(rng.standard_normal((2, 3, *src))).float()
    want = otr.affine_transform(img.double(), torch.from_numpy(th), spatial_size=dst, normalized=normalized, mode="bilinear", padding_mode="border",
                                align_corners=align, reverse_indexing=reverse, zero_centered=zero).numpy()
    m = U.affine_transform_matrix(th, src, dst, normalized, reverse, align, zero)
    for b in range(2):
        got = _sample_with_index_matrix(img[b], m, dst, "bilinear", "border", align).numpy()
        np.testing.assert_allclose(got, want[b], rtol=1e-9, atol=1e-9)
    if normalized and not reverse:   # the reference's test_forward_2d / test_forward_3d identity
        grid = F.affine_grid(torch.from_numpy(th[None, :sr]).repeat(2, 1, 1), [2, 3, *dst], align_corners=align)
        np.testing.assert_allclose(F.grid_sample(img.double(), grid, padding_mode="border", align_corners=align).numpy(), want, rtol=1e-12, atol=1e-12)


def test_affine_transform_argument_errors_follow_the_reference():
    """test_ill_affine_transform of the reference (test_affine_transform.py:264-330): the same exception types, raised before any launch."""
    from monai_b200.networks.layers import AffineTransform

    rot3 = torch.as_tensor(_ROT3, dtype=torch.float32)
    x5 = torch.arange(48.0).view(2, 1, 4, 2, 3)
    with pytest.raises(ValueError):   # image too small
        AffineTransform((3, 4, 2), padding_mode="border", align_corners=False)(torch.as_tensor([1.0, 2.0, 3.0]), rot3)
    with pytest.raises(ValueError):   # output shape too small
        AffineTransform((3, 4), padding_mode="border", align_corners=False)(x5.cuda() if torch.cuda.is_available() else _FakeCuda(x5), rot3)
    with pytest.raises(ValueError):   # incorrect affine
        AffineTransform((2, 3, 4))(x5, rot3[None, None])
    with pytest.raises(ValueError):   # batch doesn't match
        AffineTransform((2, 3, 4))(x5.cuda() if torch.cuda.is_available() else _FakeCuda(x5), rot3[None].repeat(3, 1, 1))
    with pytest.raises(RuntimeError):  # integer image
        AffineTransform((2, 3, 4), normalized=True)(x5.int(), rot3[None].repeat(2, 1, 1))
    with pytest.raises(ValueError):   # wrong affine
        AffineTransform((2, 3, 4))(x5, torch.as_tensor([[1, 0, 0, 0], [0, 0, 0, 1]]))
    with pytest.raises(RuntimeError):  # dtype doesn't match
        AffineTransform((1, 2))(torch.arange(1.0, 13.0).view(1, 1, 3, 4), torch.as_tensor([[2.0, 0.0, 0.0], [0.0, 2.0, 0.0]], dtype=torch.float64))
    with pytest.raises(TypeError):
        AffineTransform()(x5, np.eye(4))
    with pytest.raises(ValueError):
        AffineTransform(normalized=True, zero_centered=True)
    with pytest.raises(RuntimeError):  # no CPU fallback
        AffineTransform((2, 3, 4))(x5, rot3)


class _FakeCuda(torch.Tensor):
    """A CPU tensor that reports is_cuda, so that checks placed after the device test can be reached without a GPU."""

    @staticmethod
    def __new__(cls, t):
        return torch.Tensor._make_subclass(cls, t)

    @property
    def is_cuda(self):
        return True


def test_normalize_transform_and_to_norm_affine_goldens_of_the_reference_unit_test():
    """tests/networks/layers/test_affine_transform.py:27-130: every TEST_NORM_CASES / TEST_TO_NORM_AFFINE_CASES row (the zero-centred ones
    included) and the ill-formed inputs, through monai_b200.networks.utils."""
    from monai_b200.networks.utils import normalize_transform, to_norm_affine

    norm_cases = [
        ((4, 5), True, [[[0.666667, 0, -1], [0, 0.5, -1], [0, 0, 1]]], False),
        ((4, 5), True, [[[0.5, 0, 0], [0, 0.4, 0], [0, 0, 1]]], True),
        ((2, 4, 5), True, [[[2.0, 0.0, 0.0, -1.0], [0.0, 0.6666667, 0.0, -1.0], [0.0, 0.0, 0.5, -1.0], [0.0, 0.0, 0.0, 1.0]]], False),
        ((4, 5), False, [[[0.5, 0.0, -0.75], [0.0, 0.4, -0.8], [0.0, 0.0, 1.0]]], False),
        ((4, 5), False, [[[0.6666667, 0.0, 0.0], [0.0, 0.5, 0.0], [0.0, 0.0, 1.0]]], True),
        ((2, 4, 5), False, [[[1.0, 0.0, 0.0, -0.5], [0.0, 0.5, 0.0, -0.75], [0.0, 0.0, 0.4, -0.8], [0.0, 0.0, 0.0, 1.0]]], False),
    ]
    for shape, align, expected, zero in norm_cases:
        got = normalize_transform(shape, device=torch.device("cpu:0"), dtype=torch.float32, align_corners=align, zero_centered=zero)
        assert got.dtype == torch.float32 and tuple(got.shape) == (1, len(shape) + 1, len(shape) + 1)
        np.testing.assert_allclose(got.numpy(), np.asarray(expected), atol=1e-6)
    eye3, eye4 = [[[1, 0, 0], [0, 1, 0], [0, 0, 1]]], [[[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]]
    to_norm_cases = [
        (eye3, (4, 6), (5, 3), True, [[[1.3333334, 0.0, 0.33333337], [0.0, 0.4, -0.6], [0.0, 0.0, 1.0]]], False),
        (eye3, (4, 6), (5, 3), False, [[[1.25, 0.0, 0.25], [0.0, 0.5, -0.5], [0.0, 0.0, 1.0]]], False),
        (eye4, (2, 4, 6), (3, 5, 3), True, [[[2.0, 0.0, 0.0, 1.0], [0.0, 1.3333334, 0.0, 0.33333337], [0.0, 0.0, 0.4, -0.6], [0.0, 0.0, 0.0, 1.0]]], False),
        (eye4, (2, 4, 6), (3, 5, 3), False, [[[1.5, 0.0, 0.0, 0.5], [0.0, 1.25, 0.0, 0.25], [0.0, 0.0, 0.5, -0.5], [0.0, 0.0, 0.0, 1.0]]], False),
        (eye4, (2, 4, 6), (3, 5, 3), False, [[[2.0, 0.0, 0.0, 0.0], [0.0, 1.3333334, 0.0, 0.0], [0.0, 0.0, 0.4, 0.0], [0.0, 0.0, 0.0, 1.0]]], True),
    ]
    for affine, src, dst, align, expected, zero in to_norm_cases:
        got = to_norm_affine(torch.as_tensor(affine, dtype=torch.float32), src, dst, align, zero)
        np.testing.assert_allclose(got.numpy(), np.asarray(expected), atol=1e-6)
    for affine, src, dst, align in [(eye3, (3, 4, 6), (3, 5, 3), False), (eye4, (4, 6), (3, 5, 3), True),
                                    ([[[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0]]], (4, 6), (3, 5, 3), True)]:
        with pytest.raises(TypeError):
            to_norm_affine(affine, src, dst, align)
        with pytest.raises(ValueError):
            to_norm_affine(torch.as_tensor(affine, dtype=torch.float32), src, dst, align)


def _stand_in_resample(img, mat, r, out_shape, mode, padding_mode, align):
    """numerically real CPU stand-in for spatial._resample (F.grid_sample evaluates the index matrix)"""
    mode = {"bilinear": "bilinear", "linear": "bilinear", "trilinear": "bilinear", "nearest": "nearest", 1: "bilinear", 0: "nearest"}[getattr(mode, "value", mode)]
    pad = {"zeros": "zeros", "constant": "zeros", "border": "border", "nearest": "border", "reflection": "reflection", "reflect": "reflection"}[str(getattr(padding_mode, "value", padding_mode))]
    t = img.as_subclass(torch.Tensor) if type(img) is not torch.Tensor else img
    return _sample_with_index_matrix(t, np.asarray(mat, dtype=np.float64), tuple(int(s) for s in out_shape), mode, pad, bool(align)).float()


def test_host_path_with_a_stand_in_kernel_reproduces_the_real_reference_fixtures(monkeypatch, golden_dir):
    """The shipped host code of Spacing / Spacingd / RandAffined / Compose(lazy) -- including the memoised affine algebra: every
    transform object is called twice -- with only the resampling kernel replaced by a CPU evaluation of the SAME index matrix:
    values, shapes and affines of the real-reference fixtures the GPU tests use (tests/golden/transforms.npz, lazy_inverse.npz)."""
    import ast
    import os

    import monai_b200.transforms.spatial as S
    from monai_b200.data import MetaTensor
    from monai_b200.transforms import Compose, RandAffined, Spacing, Spacingd

    monkeypatch.setattr(S, "_resample", _stand_in_resample)
    g = np.load(os.path.join(golden_dir, "transforms.npz"))
    for tag in ("s0", "s1", "s2", "s3", "s4"):
        kw = ast.literal_eval(str(g[f"{tag}.kw"]))
        sp = Spacing(pixdim=tuple(g[f"{tag}.pixdim"]), **kw)
        for rep in range(2):     # the second call takes the cached algebra
            r = sp(MetaTensor(torch.from_numpy(g["img"]), affine=torch.as_tensor(g[f"{tag}.affine"])))
            assert tuple(r.shape) == g[f"{tag}.y"].shape, (tag, rep)
            np.testing.assert_allclose(r.affine.numpy(), g[f"{tag}.new_affine"], rtol=1e-9, atol=1e-9, err_msg=f"{tag} call {rep}")
            if kw.get("mode") == "nearest":
                assert (r.numpy() != g[f"{tag}.y"]).mean() < 2e-3, tag
            else:
                np.testing.assert_allclose(r.numpy(), g[f"{tag}.y"], rtol=1e-3, atol=1e-4, err_msg=f"{tag} call {rep}")
    for tag in ("r0", "r1", "r2"):
        kw = ast.literal_eval(str(g[f"{tag}.kw"]))
        t = RandAffined(keys=["image"], **kw)
        t.set_random_state(seed=0)
        r = t({"image": MetaTensor(torch.from_numpy(g["img2"]), affine=torch.eye(4))})["image"]
        assert tuple(r.shape) == g[f"{tag}.y"].shape, tag
        if kw["mode"] != "nearest":
            np.testing.assert_allclose(r.numpy(), g[f"{tag}.y"], rtol=1e-3, atol=2e-4, err_msg=tag)
        np.testing.assert_allclose(r.affine.numpy(), g[f"{tag}.new_affine"], rtol=1e-5, atol=1e-5, err_msg=tag)
    li = np.load(os.path.join(golden_dir, "lazy_inverse.npz"))
    for tag, lazy in (("eager", False), ("lazy", True)):
        c = Compose([Spacingd(keys=["image"], pixdim=(1.0, 1.0, 1.0), mode="bilinear"),
                     RandAffined(keys=["image"], prob=1.0, rotate_range=(0.2,) * 3, scale_range=(0.1,) * 3, translate_range=(5,) * 3, mode="bilinear", padding_mode="border")],
                    lazy=lazy)
        for rep in range(2):
            c.transforms[1].set_random_state(seed=0)
            y = c({"image": MetaTensor(torch.from_numpy(li["x"]), affine=torch.as_tensor(li["x_affine"]))})["image"]
            np.testing.assert_allclose(y.numpy(), li[f"{tag}.y"], rtol=1e-4, atol=1e-4, err_msg=f"{tag} call {rep}")
            np.testing.assert_allclose(np.asarray(y.affine), li[f"{tag}.affine"], atol=1e-6)


def test_grid_autograd_glue_with_oracle_backed_kernels_reproduces_the_reference_gradient_rows(monkeypatch, golden_dir):
    """The autograd Functions of grid_pull / grid_push / grid_count (networks/layers/spatial_transforms.py) with the four kernel wrappers
    replaced by the numpy restatements: the reference's test_grid_pull gradient check, all 224 rows of 1D_BP_bwd.txt, plus the 3-D backward
    fixtures of the compiled reference.  (The same assertions run against the real kernels in tests/test_gpu_zz_grid_autograd.py.)"""
    import os

    import monai_b200.networks.layers.spatial_transforms as ST
    from oracle import resample as orr

    def k_pull(x, g, b, o, extrapolate=True, channel_last=True, **kw):
        return torch.from_numpy(orr.grid_pull(x.detach().numpy(), g.detach().numpy(), b, o, extrapolate))

    def k_push(x, g, shape, b, o, extrapolate=True):
        if x is None:
            return torch.from_numpy(orr.grid_count(g.detach().numpy(), shape, b, o, extrapolate))
        return torch.from_numpy(orr.grid_push(x.detach().numpy(), g.detach().numpy(), shape, b, o, extrapolate))

    def k_grad(x, g, b, o, extrapolate=True):
        return torch.from_numpy(orr.grid_grad(x.detach().numpy(), g.detach().numpy(), b, o, extrapolate))

    monkeypatch.setattr(ST.K, "grid_pull", k_pull)
    monkeypatch.setattr(ST.K, "grid_push", k_push)
    monkeypatch.setattr(ST.K, "grid_grad", k_grad)
    fake = lambda t, g=False: _FakeCuda(t).requires_grad_(g)   # noqa: E731
    gp = np.load(os.path.join(golden_dir, "grid_push.npz"))
    rows, labels = gp["bp1d_bwd.all_rows"], gp["bp1d_bwd.all_labels"]
    for i in range(0, 224, 4):
        it, bt = str(labels[i]).split()
        for j, (input_g, grid_g) in enumerate(((True, True), (True, False), (False, True), (False, False))):
            want = rows[i + j][~np.isnan(rows[i + j])]
            x = fake(torch.arange(10, dtype=torch.float32).reshape(1, 1, 10), input_g)
            base = fake(torch.arange(20, dtype=torch.float32).reshape(1, 20, 1), grid_g)
            res = ST.grid_pull(x, base + 0.5, interpolation=it.split(".")[1], bound=bt.split(".")[1])
            grads = []
            if input_g or grid_g:
                res.sum().backward()
            if input_g:
                grads.append(x.grad.as_subclass(torch.Tensor).view(-1))
            if grid_g:
                grads.append(base.grad.as_subclass(torch.Tensor).view(-1))
            got = torch.cat(grads).numpy() if grads else np.zeros(1)
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4, err_msg=f"{labels[i]} input_g={input_g} grid_g={grid_g}")
    names = {0: "replicate", 1: "dct1", 2: "dct2", 3: "dst1", 4: "dst2", 5: "dft", 7: "zero"}
    for i in range(int(gp["n_bwd"])):
        bound, order, extrap = (int(v) for v in gp[f"b{i}.cfg"])
        kw = dict(interpolation=order, bound=names[bound], extrapolate=bool(extrap))
        tol = dict(rtol=1e-4, atol=2e-5, err_msg=f"case {i}")
        t = lambda k, g=False: fake(torch.from_numpy(gp[f"b{i}.{k}"]), g)   # noqa: E731
        pl = lambda a: a.as_subclass(torch.Tensor).numpy()   # noqa: E731
        x, grid = t("x", True), t("grid", True)
        ST.grid_pull(x, grid, **kw).backward(torch.from_numpy(gp[f"b{i}.gout"]))
        np.testing.assert_allclose(pl(x.grad), gp[f"b{i}.pull_dx"], **tol)
        np.testing.assert_allclose(pl(grid.grad), gp[f"b{i}.pull_dg"], **tol)
        xin, grid = t("xin", True), t("grid", True)
        ST.grid_push(xin, grid, (6, 5, 7), **kw).backward(torch.from_numpy(gp[f"b{i}.gvol"]))
        np.testing.assert_allclose(pl(xin.grad), gp[f"b{i}.push_dx"], **tol)
        np.testing.assert_allclose(pl(grid.grad), gp[f"b{i}.push_dg"], **tol)
        grid = t("grid", True)
        ST.grid_count(grid, (6, 5, 7), **kw).backward(torch.from_numpy(gp[f"b{i}.gcnt"]))
        np.testing.assert_allclose(pl(grid.grad), gp[f"b{i}.count_dg"], **tol)
    with pytest.raises(NotImplementedError):
        ST.grid_grad(fake(torch.zeros(1, 1, 4, 4, 4), True), fake(torch.zeros(1, 2, 2, 2, 3)))


@pytest.mark.parametrize("case", range(len(AFFINE_TRANSFORM_GOLDENS)))
def test_affine_transform_layer_with_a_stand_in_kernel_reproduces_the_reference_goldens(monkeypatch, case):
    """monai_b200.networks.layers.AffineTransform as shipped (argument handling, per-item matrices, stacking, dtype) with the resampling
    kernel replaced by its CPU stand-in: the goldens of the reference's unit test.  The GPU twin is tests/test_gpu_zz_affine_transform.py."""
    import monai_b200.transforms.spatial as S
    from monai_b200.networks.layers import AffineTransform

    monkeypatch.setattr(S, "_resample", _stand_in_resample)
    init, image, theta, call_size, expected, atol = AFFINE_TRANSFORM_GOLDENS[case]
    image = _FakeCuda(torch.as_tensor(np.asarray(image), dtype=torch.float32))
    theta = torch.as_tensor(np.asarray(theta), dtype=torch.float32)
    out = AffineTransform(**init)(image, theta, call_size)
    assert out.dtype == torch.float32
    np.testing.assert_allclose(out.as_subclass(torch.Tensor).numpy(), np.asarray(expected), atol=max(atol, 1e-4), rtol=1e-4)
