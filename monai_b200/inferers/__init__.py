from .inferer import Inferer, PatchInferer, SimpleInferer, SliceInferer, SlidingWindowInferer, SlidingWindowInfererAdapt
from .merger import AvgMerger, Merger
from .splitter import SlidingWindowSplitter, Splitter
from .utils import resample_matrix, sliding_window_inference, sliding_window_inference_resampled

__all__ = ["Inferer", "SimpleInferer", "PatchInferer", "SlidingWindowInferer", "SlidingWindowInfererAdapt", "SliceInferer", "Splitter",
           "SlidingWindowSplitter", "Merger", "AvgMerger", "sliding_window_inference", "sliding_window_inference_resampled", "resample_matrix"]
