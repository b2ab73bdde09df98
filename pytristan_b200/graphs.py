"""CUDA-graph replay of operator applications (launch-bound regime).

A single field on a BASELINE grid (config 3 at B = 1: 4 MB) is differentiated in ~10 us of GPU time
per kernel, less than the ~25 us of Python + ctypes it takes to launch one: the step is bound by the
host.  ``CapturedApply`` records the launches of ``op.apply(u, out=y)`` for a fixed set of buffers
once and replays them as one graph launch — the kernels, their arguments and their order are exactly
those of the eager path (they are captured from it), so the results are bit-identical.

Everything an operator needs on the device is built when the operator is built (packed forms,
even-odd eligibility), so ``apply`` neither allocates nor synchronises and is capturable as it is.
"""
from . import _lib

__all__ = ["CapturedApply"]


class CapturedApply:
    """``CapturedApply(ops, u, outs=None)``: the launches of ``ops[k].apply(u, out=outs[k])`` as a CUDA graph.

    ``u`` and ``outs`` are CUDA float64 tensors whose STORAGE is what the graph reads and writes: refill
    ``u`` in place (``u.copy_(...)``) and call ``replay()``.  ``ops`` may hold anything with an
    ``apply(u, out=...)`` method (operators, ``PolarLaplacian``).
    """

    def __init__(self, ops, u, outs=None):
        torch = _lib.require_cuda()
        self.ops = list(ops)
        if not isinstance(u, torch.Tensor) or not u.is_cuda or u.dtype != torch.float64:
            raise TypeError("CapturedApply needs a float64 CUDA tensor (graphs replay fixed device addresses)")
        self.u = u
        self.outs = [torch.empty_like(u) for _ in self.ops] if outs is None else list(outs)
        if len(self.outs) != len(self.ops):
            raise ValueError("CapturedApply: one output per operator")
        # warm-up on a side stream (first launches set kernel attributes, which is not capturable)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self._run()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._run()

    def _run(self):
        for op, y in zip(self.ops, self.outs):
            op.apply(self.u, out=y)

    def replay(self):
        """Launch the captured kernels on the current stream; returns the output tensors."""
        self.graph.replay()
        return self.outs
