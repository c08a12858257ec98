"""CUDA-graph capture of a whole training iteration.

Once `mean_count` is set the marcher's buffers have a fixed size, every kernel of the step goes to
the current stream through the C-ABI and nothing reads back to the host, so render -> loss ->
backward -> GradScaler -> fused Adam can be captured once and replayed: the ~60 launches of a step
cost one `cudaGraphLaunch` instead of ~60 Python/ctypes round trips.

    step = GraphedStep(fn, example_inputs)     # fn(*tensors) -> tensor(s); runs 3 eager warm-ups, then captures
    out = step(*new_inputs)                    # copies into the static inputs, replays, returns the static outputs

Limits (inherent to graphs): shapes and control flow are frozen at capture; python-side counters
(`NeRFRenderer.local_step`) do not advance during replays, so call `update_extra_state()` outside
and re-capture after it changes `mean_count`.
"""
import torch


class GraphedStep:
    def __init__(self, fn, example_inputs, warmup=3, pool=None):
        self.fn = fn
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.static_out = fn(*self.static_in)
        self.launches_per_replay = _lib.launch_count() - n0     # kernels of this library inside one replay

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
