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
