from .intensity import GaussianSmooth, GaussianSmoothd
from .lazy import apply_pending, apply_pending_transforms
from .post import Activations, Activationsd, AsDiscrete, AsDiscreted, Invertd
from .spatial import Affine, AffineGrid, RandAffine, RandAffined, RandAffineGrid, Resample, Spacing, Spacingd, SpatialResample
from .transform import Compose, MapTransform, Randomizable, RandomizableTransform, Transform
