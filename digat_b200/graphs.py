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
