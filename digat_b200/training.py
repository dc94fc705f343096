"""Training helpers: a whole training step (forward, backward kernels, gradient all-reduce, clipping, optimizer) captured
once as a CUDA graph and replayed per batch.

At the reference's batch size (64 behaviours x 5 candidates = 320 encoder rows, config.py:31,34) a step is ~300 short
kernel launches plus PyTorch's autograd and optimizer bookkeeping: eager execution is bound by the host, not by the GPU.
All shapes of a training step are fixed (no node pruning, no data-dependent launch parameters), so the step can be
captured as it is.  Requirements: fixed batch shapes, an optimizer constructed with ``capturable=True``, and a step
function without host synchronisation (``loss.item()`` belongs outside).

Data parallelism (reference trainer.py:19 wraps the model in DistributedDataParallel): ``FlatGradients`` lays every
parameter's ``.grad`` out as a view into ONE flat buffer, so the gradient exchange of a step is a single
``all_reduce(flat, AVG)`` over NCCL / NVLink -- the same result as DDP's bucketed all-reduce (mean over ranks), but one
collective without autograd hooks, which (unlike DDP's reducer) captures into the step's CUDA graph.  21.2 MB at L=3: about
0.1 ms on NVSwitch, so overlapping it with the backward (what DDP's buckets are for) would buy nothing here."""
import torch


class FlatGradients:
    """Gives every parameter a persistent ``.grad`` that is a view into one flat fp32 buffer.

    ``zero()`` replaces ``optimizer.zero_grad()`` (the views must stay in place: autograd then accumulates into them in
    place), ``all_reduce_mean()`` averages the buffer over the process group (no-op for a single process)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if dist.get_backend(group) == 'nccl':
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
            else:                                           # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                self.flat.div_(dist.get_world_size(group))


class FlatAdam:
    """Clip-by-global-norm + Adam (reference trainer.py:98-105: ``clip_grad_norm_`` then ``optim.Adam.step``) over flat buffers,
    two kernel launches per step (digat_grad_sumsq, digat_adam_clip_step) instead of PyTorch's per-parameter kernels.

    Parameters are re-pointed (``p.data``) into ONE flat fp32 buffer, their ``.grad`` into another (FlatGradients), and the two
    moment buffers are flat too.  The step counter lives on the device, so the whole step captures into a CUDA graph.
    ``weight_decay`` is torch.optim.Adam's L2 term and applies to every parameter (the reference's default is 0, config.py:36;
    its bias / LayerNorm exemption list matches no parameter of this encoder except the biases)."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=1.0, modules=()):
        """modules: nn.Modules whose encoders keep packed inference copies of the weights (DIGAT / MSA / CNN ``invalidate_packed``):
        the update kernel writes the parameters in place without bumping ``param._version``, so every step invalidates them."""
        from . import _lib
        self._packed_owners = [m for mod in modules for m in mod.modules() if hasattr(m, 'invalidate_packed')]
        self._lib = _lib
        self.grads = FlatGradients(params)
        self.params = self.grads.params
        dev = self.params[0].device
        if dev.type != 'cuda':
            raise RuntimeError('FlatAdam needs CUDA parameters (digat_b200 has no CPU fallback)')
        _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
        n = self.grads.flat.numel()
        self.flat_params = torch.empty(n, dtype=torch.float32, device=dev)
        off = 0
        with torch.no_grad():
            for p in self.params:
                view = self.flat_params[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                off += p.numel()
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        self.partials = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm

    def zero_grad(self):
        self.grads.zero()

    def all_reduce_mean(self, group=None):
        self.grads.all_reduce_mean(group)

    def step(self):
        stream = torch.cuda.current_stream().cuda_stream
        g, n = self.grads.flat, self.grads.flat.numel()
        self._lib.call('digat_grad_sumsq', g.data_ptr(), n, self.partials.data_ptr(), self.step_count.data_ptr(), stream)
        self._lib.call('digat_adam_clip_step', self.flat_params.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                       self.exp_avg_sq.data_ptr(), n, self.partials.data_ptr(), self.step_count.data_ptr(),
                       self.grad_norm.data_ptr(), float(self.max_norm or 0.0), float(self.lr), float(self.betas[0]),
                       float(self.betas[1]), float(self.eps), float(self.weight_decay), stream)
        for m in self._packed_owners:
            m.invalidate_packed()


def broadcast_parameters(module, src=0, group=None):
    """What DDP does at construction: every rank starts from rank ``src``'s parameters and buffers."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)
        for m in module.modules():
            if hasattr(m, 'invalidate_packed'):
                m.invalidate_packed()


class GraphedTrainStep:
    def __init__(self, step_fn, example_inputs, warmup: int = 3, modules=(), distributed: bool = False):
        """step_fn(*inputs) -> loss tensor; it must do zero_grad / backward / optimizer.step itself.
        NOTE: the ``warmup`` eager calls are real training steps on ``example_inputs`` (the capture itself executes nothing).
        modules: nn.Modules whose DIGAT encoders keep a packed copy of the weights for inference -- a replayed optimizer
        step does not bump ``param._version``, so every replay invalidates those copies (DIGAT.invalidate_packed).
        distributed: the step contains NCCL collectives (FlatGradients.all_reduce_mean): capture in thread-local error mode
        (NCCL's watchdog thread polls events while the capture is open)."""
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
        kw = {'capture_error_mode': 'thread_local'} if distributed else {}
        with torch.cuda.graph(self.graph, **kw):
            self.static_loss = step_fn(*self.static_inputs)

    def __call__(self, *inputs):
        for dst, src in zip(self.static_inputs, inputs):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for m in self._encoders:
            m.invalidate_packed()
        return self.static_loss
