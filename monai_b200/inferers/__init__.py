from .inferer import Inferer, SimpleInferer, SlidingWindowInferer, SlidingWindowInfererAdapt
from .utils import sliding_window_inference

__all__ = ["Inferer", "SimpleInferer", "SlidingWindowInferer", "SlidingWindowInfererAdapt", "sliding_window_inference"]
