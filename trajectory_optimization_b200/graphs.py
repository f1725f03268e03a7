"""CUDA-graph capture of one optimisation step (the step is launch-bound once the objective is fused).

`GraphedStep(model, optimizer)` captures `optimizer.zero_grad(); loss = model(); loss.backward(); optimizer.step()`
— what the reference's loops do every iteration (src/pose_optimization.py:129-137,
src/trajectory_optimization.py:106-116) — into one `torch.cuda.CUDAGraph` and replays it: one launch per step
instead of ~60 (ModelPose) to ~150 (ModelTraj with its regularisers).  Every kernel of libcovb200.so is
capturable: the library launches only on the stream it is given, never allocates and never synchronises.

The optimiser must keep its step counter on the device (`torch.optim.Adam(..., capturable=True)`); parameters must
stay the same objects between replays (update them in place).  A captured graph holds the ADDRESSES of the buffers the
kernels read — for `ModelTraj` that is the Morton-ordered copy of the cloud, its permutation and its tile boxes, not
`model.points` itself — so a new cloud of the same size goes in through `model.refresh_points_(new_points)`, which
re-orders into those same buffers; a cloud of another size needs `model.set_points()` and a new capture.
"""
import torch


class GraphedStep:
    def __init__(self, model, optimizer, forward_kwargs=None, warmup=3, stream=None):
        if not torch.cuda.is_available():
            raise RuntimeError("GraphedStep needs a CUDA device")
        self.model, self.optimizer = model, optimizer
        self.kwargs = dict(forward_kwargs or {})
        # A model that already ran eagerly on the default stream keeps its last autograd graph alive (ModelTraj.loss[...],
        # ModelPose._total); the AccumulateGrad nodes of that graph are bound to the default stream and would make the
        # capture depend on it ("operation would make the legacy stream depend on a capturing blocking stream").
        if hasattr(model, "detach_state"):
            model.detach_state()
        optimizer.zero_grad(set_to_none=True)
        side = stream or torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the default stream, as capture requires
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = model(**self.kwargs)
            self.loss.backward()
            optimizer.step()

    def _eager(self):
        self.optimizer.zero_grad(set_to_none=True)
        loss = self.model(**self.kwargs)
        loss.backward()
        self.optimizer.step()
        return loss

    def step(self):
        """Replay the captured step; returns the (static) loss tensor of this step."""
        self.graph.replay()
        return self.loss


class GraphedCall:
    """Capture an arbitrary no-argument callable (e.g. objective + backward without an optimiser) and replay it."""

    def __init__(self, fn, warmup=3):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out
