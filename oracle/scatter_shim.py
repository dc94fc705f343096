"""TEST INFRASTRUCTURE ONLY -- pure-torch restatement of the two ``torch_scatter`` ops DIGAT calls.

``torch_scatter`` (pins: ``torch_scatter==2.0.9`` in reference README.md:12, ``torch-scatter==2.1.1`` in
install_dependencies.sh:16) is a third-party dependency that is NOT vendored under /root/reference and is not
installed in this image.  Its published algorithm for the two call sites graphEncoders.py:129-130 is restated here:

* ``scatter_sum(src, index, dim, dim_size)``  = ``zeros(dim_size).scatter_add_(dim, broadcast(index), src)``
  (torch_scatter/scatter.py ``scatter_sum``); empty segments stay 0; on CPU the adds happen in source order.
* ``scatter_softmax(src, index, dim)``        = ``src - scatter_max(src)[index]`` -> ``exp`` ->
  ``/ scatter_sum(exp)[index]`` (torch_scatter/composite/softmax.py).  Release 2.0.x adds ``eps=1e-12`` to the
  denominator, which is a no-op in fp32/fp64 because the denominator is >= 1 (the max element contributes exp(0)).

Parity: UNPINNED by the reference (it has no tests / golden vectors for this boundary); pinned here only against
the mathematical definition (tests/test_oracle.py checks it against a per-segment ``torch.softmax``).
"""
import torch


def _broadcast(index: torch.Tensor, src: torch.Tensor, dim: int) -> torch.Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    while index.dim() < src.dim():
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter_max_values(src, index, dim, dim_size):
    index = _broadcast(index, src, dim)
    size = list(src.size())
    size[dim] = dim_size
    out = torch.full(size, float('-inf'), dtype=src.dtype, device=src.device)
    return out.scatter_reduce_(dim, index, src, reduce='amax', include_self=True)


def scatter_softmax(src, index, dim=-1, dim_size=None):
    if not torch.is_floating_point(src):
        raise ValueError('`scatter_softmax` can only be computed over tensors with floating point data types.')
    index = _broadcast(index, src, dim)
    if dim_size is None:
        dim_size = int(index.max()) + 1
    max_per_index = scatter_max_values(src, index, dim, dim_size).gather(dim, index)
    recentered = (src - max_per_index).exp_()
    sums = scatter_sum(recentered, index, dim, dim_size=dim_size).gather(dim, index)
    return recentered.div(sums)
