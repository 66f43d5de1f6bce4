from .inferer import Inferer, SimpleInferer, SlidingWindowInferer, SlidingWindowInfererAdapt, SliceInferer
from .utils import sliding_window_inference

__all__ = ["Inferer", "SimpleInferer", "SlidingWindowInferer", "SlidingWindowInfererAdapt", "SliceInferer", "sliding_window_inference"]
