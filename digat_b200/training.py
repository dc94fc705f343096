"""Training helper: a whole training step (forward, backward kernels, gradient clipping, optimizer) captured once as a
CUDA graph and replayed per batch.

At the reference's batch size (64 behaviours x 5 candidates = 320 encoder rows, config.py:31,34) a step is ~300 short
kernel launches plus PyTorch's autograd and optimizer bookkeeping: eager execution is bound by the host, not by the GPU.
All shapes of a training step are fixed (no node pruning, no data-dependent launch parameters), so the step can be
captured as it is.  Requirements: fixed batch shapes, an optimizer constructed with ``capturable=True``, and a step
function without host synchronisation (``loss.item()`` belongs outside)."""
import torch


class GraphedTrainStep:
    def __init__(self, step_fn, example_inputs, warmup: int = 3, modules=()):
        """step_fn(*inputs) -> loss tensor; it must do zero_grad / backward / optimizer.step itself.
        NOTE: the ``warmup`` eager calls and the capture itself are real training steps on ``example_inputs``.
        modules: nn.Modules whose DIGAT encoders keep a packed copy of the weights for inference -- a replayed optimizer
        step does not bump ``param._version``, so every replay invalidates those copies (DIGAT.invalidate_packed)."""
        self._encoders = [m for mod in modules for m in mod.modules() if hasattr(m, 'invalidate_packed')]
        self.static_inputs = [x.clone() for x in example_inputs]
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):                      # warm up on a side stream (PyTorch's capture recipe)
            for _ in range(warmup):
                step_fn(*self.static_inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_loss = step_fn(*self.static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for m in self._encoders:
            m.invalidate_packed()
        return self.static_loss
