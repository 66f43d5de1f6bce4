from .intensity import GaussianSmooth, GaussianSmoothd
from .spatial import Affine, AffineGrid, RandAffine, RandAffined, RandAffineGrid, Resample, Spacing, Spacingd, SpatialResample
from .transform import Compose, MapTransform, Randomizable, RandomizableTransform, Transform
