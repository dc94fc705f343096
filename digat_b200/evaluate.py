"""Ranking + MIND metrics for scored pairs (host side, numpy).  Mirrors the observable behaviour of reference
util.py:70-80 (per-impression rank lists from a STABLE descending sort: ties keep candidate order) and
evaluate.py:7-89 (AUC / MRR / nDCG@5 / nDCG@10 on score = 1/rank, averaged over impressions).
tests/ compare it with the oracle's restatement and with the golden metrics of the unmodified evaluate.scoring."""
import numpy as np


def rank_lists(scores: np.ndarray, impression_of_pair: np.ndarray):
    scores = np.asarray(scores, dtype=np.float64)       # util.py:70 `scores.tolist()` -> python floats
    imp = np.asarray(impression_of_pair)
    n_imp = int(imp[-1]) + 1 if len(imp) else 0
    counts = np.bincount(imp, minlength=n_imp)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    out = []
    for k in range(n_imp):
        s = scores[starts[k]:starts[k] + counts[k]]
        order = np.argsort(-s, kind='stable')
        r = np.empty(len(s), dtype=np.int64)
        r[order] = np.arange(1, len(s) + 1)
        out.append(r.tolist())
    return out


def _auc(y_true, y_score):
    # rank-sum AUC with midranks for ties (what sklearn.metrics.roc_auc_score computes for binary labels)
    order = np.argsort(y_score, kind='mergesort')
    s = y_score[order]
    ranks = np.empty(len(s), dtype=np.float64)
    i = 0
    while i < len(s):
        j = i
        while j + 1 < len(s) and s[j + 1] == s[i]:
            j += 1
        ranks[i:j + 1] = 0.5 * (i + j) + 1.0
        i = j + 1
    r = np.empty(len(s), dtype=np.float64)
    r[order] = ranks
    pos = y_true > 0
    n_pos, n_neg = int(pos.sum()), int((~pos).sum())
    if n_pos == 0 or n_neg == 0:
        raise ValueError('Only one class present in y_true. ROC AUC score is not defined in that case.')
    return (r[pos].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg)


def _dcg(y_true, y_score, k):
    order = np.argsort(y_score)[::-1]
    y = np.take(y_true, order[:k])
    return np.sum((2 ** y - 1) / np.log2(np.arange(len(y)) + 2))


def metrics(ranks, labels_per_impression):
    aucs, mrrs, n5, n10 = [], [], [], []
    for r, lab in zip(ranks, labels_per_impression):
        if len(lab) == 0:
            continue
        y_true = np.asarray(lab, dtype='float32')
        y_score = np.array([1.0 / x for x in r])
        aucs.append(_auc(y_true, y_score))
        order = np.argsort(y_score)[::-1]
        yt = np.take(y_true, order)
        mrrs.append(np.sum(yt / (np.arange(len(yt)) + 1)) / np.sum(yt))
        n5.append(_dcg(y_true, y_score, 5) / _dcg(y_true, y_true, 5))
        n10.append(_dcg(y_true, y_score, 10) / _dcg(y_true, y_true, 10))
    return float(np.mean(aucs)), float(np.mean(mrrs)), float(np.mean(n5)), float(np.mean(n10))


# ----------------------------------------------------------------------------------------------- device versions
def impression_offsets(impression_of_pair: np.ndarray):
    """[P] non-decreasing impression id per pair -> [n_imp+1] int64 offsets into the ordered pair list."""
    imp = np.asarray(impression_of_pair)
    n_imp = int(imp[-1]) + 1 if len(imp) else 0
    off = np.zeros(n_imp + 1, dtype=np.int64)
    np.cumsum(np.bincount(imp, minlength=n_imp), out=off[1:])
    return off


def rank_pairs_device(scores, offsets):
    """scores [P] fp32 CUDA tensor, offsets [n_imp+1] int64 CUDA tensor -> ranks [P] int32 CUDA tensor: the rank lists of
    ``rank_lists`` (reference util.py:70-80) concatenated in pair order, computed by rank_impressions_kernel."""
    import torch
    from . import _lib
    dev = scores.device
    _lib.require_device(dev.index if dev.index is not None else torch.cuda.current_device())
    assert scores.dtype == torch.float32 and offsets.dtype == torch.int64 and scores.is_contiguous()
    ranks = torch.empty(scores.shape[0], dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call('digat_rank_impressions', scores.data_ptr(), offsets.data_ptr(), ranks.data_ptr(),
                  offsets.shape[0] - 1, torch.cuda.current_stream().cuda_stream)
    return ranks


def metrics_device(ranks, labels, offsets, on_single_class='raise'):
    """ranks [P] int32, labels [P] (0/1, any integer/bool dtype), offsets [n_imp+1] int64, all CUDA tensors ->
    (AUC, MRR, nDCG@5, nDCG@10) averaged over the non-empty impressions, as reference evaluate.py:66-89.
    Per-impression values come from impression_metrics_kernel in double; the mean is taken on the host in impression
    order.  Raises ValueError if an impression has a single class (sklearn's roc_auc_score raises there; MIND impressions
    always hold both) unless on_single_class='skip' (synthetic corpora with one-candidate impressions)."""
    import torch
    from . import _lib
    dev = ranks.device
    n_imp = offsets.shape[0] - 1
    lab = labels.to(torch.uint8).contiguous()
    out = torch.empty((n_imp, 4), dtype=torch.float64, device=dev)
    valid = torch.empty(n_imp, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        _lib.call('digat_impression_metrics', ranks.data_ptr(), lab.data_ptr(), offsets.data_ptr(), out.data_ptr(),
                  valid.data_ptr(), n_imp, torch.cuda.current_stream().cuda_stream)
    v = valid.cpu().numpy()
    if (v == 2).any() and on_single_class != 'skip':
        raise ValueError('Only one class present in y_true. ROC AUC score is not defined in that case.')
    o = out.cpu().numpy()[v == 1]
    return tuple(float(np.mean(o[:, k])) for k in range(4))
