"""Host-side (numpy, vectorised) builders for the integer/bool graph inputs of the DIGAT encoder.

``build_user_graphs`` restates the rule of reference MIND_corpus.py:146-176 without its O(H^2) Python loops;
``oracle/digat_oracle.py::user_graph_loops`` keeps the literal loops and tests/ compare the two bit-exactly.
"""
import numpy as np


def build_user_graphs(history_category: np.ndarray, history_len: np.ndarray, H: int, C: int):
    """history_category [N,H] int (category id of history slot t, any value for t >= history_len[n]).

    Returns (user_graph [N,H+C,H+C] bool, category_mask [N,C+1] bool, category_indices [N,H] int64).
    Node layout: 0..H-1 history news, H..H+C-1 topic nodes (reference graphEncoders.py:14, MIND_corpus.py:136).
    """
    N = history_category.shape[0]
    n_u = H + C
    valid = np.arange(H)[None, :] < history_len[:, None]                      # [N,H]
    cidx = np.where(valid, history_category, C).astype(np.int64)              # padding bucket C (MIND_corpus.py:149)
    onehot = np.zeros((N, H, C + 1), dtype=bool)
    np.put_along_axis(onehot, cidx[:, :, None], True, axis=2)
    onehot_c = onehot[:, :, :C] & valid[:, :, None]                           # [N,H,C] slot t has category c
    cmask = np.zeros((N, C + 1), dtype=bool)
    cmask[:, :C] = onehot_c.any(axis=1)                                       # bucket C stays 0 (MIND_corpus.py:147)
    g = np.zeros((N, n_u, n_u), dtype=bool)
    same = (cidx[:, :, None] == cidx[:, None, :]) & valid[:, :, None] & valid[:, None, :]
    g[:, :H, :H] = same                                                       # news-news edges (:168-170) + diag
    g[:, :H, H:] = onehot_c                                                   # news-topic edges (:163-164)
    g[:, H:, :H] = onehot_c.transpose(0, 2, 1)
    present = cmask[:, :C]
    tt = present[:, :, None] & present[:, None, :]                            # topic-topic edges between the
    g[:, H:, H:] = tt & ~np.eye(C, dtype=bool)[None]                          # distinct categories present (:171-173)
    idx = np.arange(n_u)
    g[:, idx, idx] = True                                                     # identity (:145)
    return g, cmask, cidx


# ----------------------------------------------------------------------------------------------- device builders
def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def build_user_graphs_device(history_category, history_len, H: int, C: int, check: bool = True):
    """CUDA version of ``build_user_graphs`` (kernel user_graph_kernel, csrc/builders.cuh): the same three outputs as
    torch tensors on the device of ``history_category`` ([N,H] integer CUDA tensor; ``history_len`` [N]).

    The reference builds these per behaviour with O(H^2) Python loops (MIND_corpus.py:143-176); a category outside
    [0,C) or a length outside [0,H] raises (the reference raises IndexError)."""
    import torch
    from . import _lib
    dev = history_category.device
    _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
    N = history_category.shape[0]
    assert history_category.shape == (N, H) and history_len.shape == (N,)
    cat = history_category.to(torch.int32).contiguous()
    ln = history_len.to(torch.int32).contiguous()
    n_u = H + C
    g = torch.empty((N, n_u, n_u), dtype=torch.uint8, device=dev)
    cmask = torch.empty((N, C + 1), dtype=torch.uint8, device=dev)
    cidx = torch.empty((N, H), dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call('digat_build_user_graphs', cat.data_ptr(), ln.data_ptr(), g.data_ptr(), cmask.data_ptr(),
                  cidx.data_ptr(), N, H, C, err.data_ptr(), _stream())
    if check and int(err.item()) != 0:
        raise IndexError('build_user_graphs_device: category outside [0,%d) or history length outside [0,%d]' % (C, H))
    return g.view(torch.bool), cmask.view(torch.bool), cidx


def similar_to_csr(similar, n_news: int):
    """``similar[k]`` = [(news_index, cos_similarity), ...] (most similar first; the layout of the reference's
    similarity dict after id mapping, construct_SAG.py:466-470) -> CSR arrays (offsets int64, idx int32, cos float64)."""
    counts = np.array([len(similar[k]) if k < len(similar) and similar[k] is not None else 0 for k in range(n_news)],
                      dtype=np.int64)
    off = np.zeros(n_news + 1, dtype=np.int64)
    np.cumsum(counts, out=off[1:])
    idx = np.zeros(int(off[-1]), dtype=np.int32)
    cos = np.zeros(int(off[-1]), dtype=np.float64)
    for k in range(n_news):
        if counts[k]:
            row = similar[k]
            idx[off[k]:off[k + 1]] = [int(r[0]) for r in row]
            cos[off[k]:off[k + 1]] = [float(r[1]) for r in row]
    return off, idx, cos


def sag_bfs_device(sim_off, sim_idx, sim_cos, n_news: int, top_M: int, hop: int, n_nodes: int, threshold: float,
                   device='cuda', check: bool = True):
    """CUDA version of the reference's ``generate_news_graph`` (construct_SAG.py:449-485) on CSR similar-news lists
    (numpy arrays or tensors).  Returns (news_node_ID int32 [N,n], news_graph bool [N,n,n], news_graph_mask bool [N,n])
    on the device; bit-identical to the reference (tests/test_gpu_builders.py)."""
    import torch
    from . import _lib
    dev = torch.device(device)
    _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
    t = lambda a, dt: (torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a).to(dev, dt).contiguous()
    off, idx, cos = t(sim_off, torch.int64), t(sim_idx, torch.int32), t(sim_cos, torch.float64)
    assert off.shape[0] == n_news + 1
    node = torch.empty((n_news, n_nodes), dtype=torch.int32, device=dev)
    g = torch.empty((n_news, n_nodes, n_nodes), dtype=torch.uint8, device=dev)
    mask = torch.empty((n_news, n_nodes), dtype=torch.uint8, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call('digat_sag_bfs', off.data_ptr(), idx.data_ptr(), cos.data_ptr(), node.data_ptr(), g.data_ptr(),
                  mask.data_ptr(), n_news, top_M, hop, n_nodes, float(threshold), err.data_ptr(), _stream())
    if check:
        e = int(err.item())
        if e & 1:
            raise IndexError('sag_bfs_device: neighbour id outside [0,%d)' % n_news)
        if e & 2:
            raise IndexError('sag_bfs_device: a graph needs more than %d nodes' % n_nodes)
    return node, g.view(torch.bool), mask.view(torch.bool)
