"""CPU harness (no GPU needed): the C4 pipeline (Spacingd + RandAffined + GaussianSmoothd) with the kernel wrappers stubbed by empty
outputs and the is_cuda checks patched out -- what is left is the pure HOST cost per volume (affine algebra, metadata, wrapper glue),
which bounds the C4 step.    python profiles/run_c4_host_stub.py [lazy]"""
import cProfile, os, pstats, sys, time, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import monai_b200._kernels as K, monai_b200._lib as L
import importlib.util, types
def load_patched(modname, path):
    src=open(path).read().replace("not img.is_cuda","False")
    spec=importlib.util.spec_from_file_location(modname, path)
    mod=importlib.util.module_from_spec(spec); sys.modules[modname]=mod
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod
import monai_b200.transforms  # package first (loads the originals)
SP=load_patched("monai_b200.transforms.spatial",os.path.join(ROOT, "monai_b200/transforms/spatial.py"))
IT=load_patched("monai_b200.transforms.intensity",os.path.join(ROOT, "monai_b200/transforms/intensity.py"))
import monai_b200.transforms as TR
for m in (SP, IT):
    for n in getattr(m,"__all__",[]): setattr(TR, n, getattr(m,n))
K.resample_affine = lambda src, out_shape, mat, interp, pad, align, out_dtype=torch.float32: torch.empty((src.shape[0], *out_shape), dtype=out_dtype)
K.separable_filter3d = lambda src, taps: torch.empty_like(src)
L.require_cuda = lambda *a: None
class FakeCuda(torch.Tensor):
    @property
    def is_cuda(self): return True
import bench
from monai_b200.data import MetaTensor
lazy = len(sys.argv) > 1 and sys.argv[1] == "lazy"
pipe = bench._transform_pipeline(lazy=lazy)
aff = torch.diag(torch.tensor([1.25, 1.25, 1.25, 1.0], dtype=torch.float64))
vols = [torch.rand((1, 16, 16, 16)) for _ in range(4)]
def mk(v):
    return MetaTensor(v, affine=aff)
try:
    for v in vols: pipe({"image": mk(v)})
except Exception as e:
    import traceback; traceback.print_exc(); sys.exit(1)
t0=time.perf_counter()
N=50
for _ in range(N):
    for v in vols: y = pipe({"image": mk(v)})["image"]
dt=(time.perf_counter()-t0)/(N*4)
print(f"host time per volume: {dt*1e3:.3f} ms  (x32 = {dt*32e3:.1f} ms)")
pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    for v in vols: y = pipe({"image": mk(v)})["image"]
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(30)
