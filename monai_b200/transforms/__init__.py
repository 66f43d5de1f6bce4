from .intensity import GaussianSmooth, GaussianSmoothd
from .post import Activations, Activationsd, AsDiscrete, AsDiscreted
from .spatial import Affine, AffineGrid, RandAffine, RandAffined, RandAffineGrid, Resample, Spacing, Spacingd, SpatialResample
from .transform import Compose, MapTransform, Randomizable, RandomizableTransform, Transform
