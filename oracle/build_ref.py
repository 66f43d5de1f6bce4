"""TEST INFRASTRUCTURE -- builds the REFERENCE's own native resampler so the oracle can be checked against it.

The reference's only native module on this path is monai._C (monai/csrc/ext.cpp:21-75).  Its CPU part compiles from its own
sources where they lie under /root/reference/monai/csrc (no reference build system, no copies of reference sources): this
script hands the .cpp files to torch.utils.cpp_extension and writes ONLY into oracle/_ref/ (git-ignored, shipped to the GPU
box with the snapshot).  `load()` imports the built module on any box that has the .so (no /root/reference needed).

    python oracle/build_ref.py          # ~70 s, single-threaded (libgomp is not linkable in this image: no -fopenmp)
"""
from __future__ import annotations

import glob
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
NAME = "monai_C_ref"
REF = "/root/reference/monai/csrc"


def so_path() -> str | None:
    hits = sorted(glob.glob(os.path.join(OUT, NAME + "*.so")))
    return hits[0] if hits else None


def build(verbose: bool = False) -> str | None:
    """Compile the reference csrc (CPU sources only) into oracle/_ref/; returns the .so path (None if the reference is absent)."""
    if so_path():
        return so_path()
    if not os.path.isdir(REF):
        return None
    import torch
    from torch.utils.cpp_extension import load

    os.makedirs(OUT, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(REF, "**", "*.cpp"), recursive=True))
    ver = torch.__version__.split("+")[0].split(".")
    tv = int(ver[0]) * 10000 + int(ver[1]) * 100 + int(ver[2])
    load(name=NAME, sources=srcs, extra_include_paths=[REF], extra_cflags=["-DAT_PARALLEL_OPENMP=1", f"-DMONAI_TORCH_VERSION={tv}", "-O2"],
         build_directory=OUT, verbose=verbose, is_python_module=True)
    return so_path()


def load():
    """Import the built reference module (grid_pull, grid_push, ...); None when it has not been built."""
    p = so_path()
    if p is None:
        return None
    import torch  # noqa: F401  (the extension links against libtorch)

    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
