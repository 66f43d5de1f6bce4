"""Splitter cases shared by the fixture generator (make_golden.py, needs the reference) and the tests (which must not)."""

PATCH_CASES = {
    # name: (input shape, splitter kwargs)
    "p2d": ((2, 1, 7, 9), dict(patch_size=(4, 4), overlap=0.5, offset=(-1, 0), pad_mode="constant", pad_value=3)),
    "p3d_nopad": ((1, 2, 10, 11, 9), dict(patch_size=(4, 5, 3), overlap=(1, 2, 0), offset=0, pad_mode=None)),
    "p3d_rep": ((1, 1, 10, 12, 9), dict(patch_size=4, overlap=0.25, offset=(0, -2, 1), pad_mode="replicate")),
}
