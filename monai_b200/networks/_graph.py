"""CUDA-graph replay of a fixed-shape forward pass (shared by the tensor-core networks).

One forward of these networks is hundreds of short kernel launches; issuing them from Python through ctypes costs
more than the kernels themselves once they are fast.  The launch sequence for a given input shape is therefore
captured once into a CUDA graph and replayed.  Graphs are keyed by (shape, dtype, device) and invalidated when any
parameter changes (`load_state_dict`, `.half()`).  MONAI_B200_GRAPH=0 disables the mechanism.
"""
from __future__ import annotations

import os

import torch

from .. import _kernels as K
from .. import _lib as L


class GraphedForward:
    def _graph_init(self) -> None:
        self._graph_enabled = os.environ.get("MONAI_B200_GRAPH", "1") != "0"
        self._graphs: dict = {}

    def _graph_ok(self) -> bool:
        return self._graph_enabled and not K._Prof.on and not torch.cuda.is_current_stream_capturing()

    def _forward_graphed(self, x_in: torch.Tensor, impl) -> torch.Tensor:
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        key = (tuple(x_in.shape), x_in.dtype, x_in.device, impl.__name__)
        ent = self._graphs.get(key)
        if ent is None or ent["sig"] != sig:
            static_in = x_in.detach().clone().contiguous()
            impl(static_in)  # eager warm-up: packs weights, sets kernel attributes, fills the plan caches
            torch.cuda.synchronize(x_in.device)
            graph = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            # thread_local: other threads (e.g. the NCCL watchdog) may issue CUDA calls while this thread captures
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                static_out = impl(static_in)
            n_kernels = L.launch_count() - n0
            if len(self._graphs) >= 4:  # bound the private memory pools kept alive
                self._graphs.pop(next(iter(self._graphs)))
            ent = {"sig": sig, "graph": graph, "inp": static_in, "out": static_out, "n_kernels": n_kernels}
            self._graphs[key] = ent
        ent["inp"].copy_(x_in)
        ent["graph"].replay()
        L.add_replayed_launches(ent["n_kernels"])
        return ent["out"].clone()
