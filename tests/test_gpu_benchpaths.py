"""Numerical parity ON THE CODE PATHS bench.py TIMES (VERDICT r1 "what's weak" 1): the small golden cases (2-6 rows) run
on the exact-fp32 CUDA-core GEMM without pruning, so here

* every golden case of the unmodified reference is replicated x64: all projections then take the tcgen05 3xTF32 GEMM
  (row-scattered, persistent for the larger ones), node pruning and the multi-graph edge-driven kernel are on -- this is
  the only tensor-core coverage of graph_depth=7 (``wide_n8_L7``); gate 1e-5 against the reference's fp32 outputs;
* each bench workload (BASELINE.json configs[1], the reference's argparse default, configs[3]) is scored at the bench's
  own batch size (4096 pairs per step, two steps through the pipelined drivers) on the full-size synthetic corpus,
  through ``score_resident_batches`` AND ``score_host_batches``; 256 sampled pairs are compared with the CPU oracle
  (reference graphEncoders.py:189-198, util.py:56-69).

Errors are reported as max-norm, per-element (stated floor), conditioned on the dot product's terms, and in ULPs."""
import numpy as np
import pytest
import torch

from oracle import digat_oracle as O
from tests.helpers import CASES, case_inputs, check_hashes, err_stats, fmt_stats, load_golden, rel_err

pytestmark = pytest.mark.gpu
ORDER = ('news_graph_embeddings', 'news_graph', 'news_graph_mask', 'user_news_embedding', 'user_graph',
         'user_category_mask', 'user_category_indices')
LOGIT_TOL = 1e-5          # north star: fp32 click logits within 1e-5 relative (max-norm, as in test_gpu_parity.py)
COND_TOL = 1e-5           # per-logit error relative to sum_d |c_n[d] * c_u[d]| (what fp32 rounding of that dot acts on)
REPLICAS = 64


@pytest.mark.parametrize('name', list(CASES))
def test_golden_cases_replicated_take_the_tensor_core_path(name):
    from digat_b200 import _lib
    from digat_b200.graphEncoders import DIGAT, TENSOR_CORE_MIN_ROWS
    from digat_b200.model import logits
    cfg, sd, corpus, batch = case_inputs(name)
    z, meta = load_golden(name)
    check_hashes(meta, sd, batch)
    m = DIGAT(cfg, 400)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    B = batch['news_graph'].shape[0]
    bb = {k: v.cuda().repeat(REPLICAS, *([1] * (v.dim() - 1))) for k, v in batch.items()}
    assert B * REPLICAS * cfg.news_graph_size >= TENSOR_CORE_MIN_ROWS      # even the news-graph GEMMs are on tcgen05
    args = [bb[k] for k in ORDER]
    prof = _lib.start_profile()
    with torch.no_grad():
        c_n0 = m.compute_news_graph_context(bb['news_graph_embeddings'], bb['news_graph_mask'])
        cn, cu = m.inference(*args, c_n0)
        fn, fu = m.forward(*args)
        lg = logits(cn, cu)
        lf = logits(fn, fu)
    torch.cuda.synchronize()
    rec = _lib.stop_profile()
    m.check_index_errors()
    names = [r[0] for r in rec]
    assert 'digat_linear_tf32x3' in names, 'the replicated batch did not reach the tensor-core GEMM'
    assert any(r[0] == 'digat_linear_tf32x3' and r[1][16] for r in rec), 'no row-scattered (pruned) projection ran'
    assert any(r[0] == 'digat_graph_layer_fwd' and r[1][19] for r in rec), 'the layer kernel ran without row_active'
    for k, v in (('logits', lg), ('news_ctx', cn), ('user_ctx', cu), ('fwd_news_ctx', fn), ('fwd_user_ctx', fu), ('c_n0', c_n0)):
        got = v.cpu().numpy()
        ref32, ref64 = z['ref32_' + k], z['ref64_' + k]
        for r in range(REPLICAS):                                           # every replica, not only the first
            e = rel_err(got[r * B:(r + 1) * B], ref32)
            assert e < LOGIT_TOL, '%s/%s replica %d: rel err vs fp32 reference %.3e' % (name, k, r, e)
        if k == 'logits':
            terms = np.abs(z['ref64_news_ctx'] * z['ref64_user_ctx']).sum(1)
            s32, s64 = err_stats(got[:B], ref32, terms), err_stats(got[:B], ref64, terms)
            sref = err_stats(ref32, ref64, terms)
            print('\n%s x%d logits vs reference fp32: %s\n    vs fp64: %s\n    reference fp32 vs fp64: %s'
                  % (name, REPLICAS, fmt_stats(s32), fmt_stats(s64), fmt_stats(sref)))
            assert s32['cond'] < COND_TOL and s64['cond'] < COND_TOL
    assert rel_err(lf.cpu().numpy()[:B], z['ref32_logits']) < LOGIT_TOL


def _oracle_logits(sd, corpus, ids, dtype):
    from digat_b200 import synth
    P = O.cast_params(sd, dtype)
    out, ctx = [], []
    with torch.no_grad():
        for s in range(0, len(ids), 64):
            b = synth.make_batch(corpus, ids[s:s + 64])
            a = [b[k].to(dtype) if b[k].is_floating_point() else b[k] for k in ORDER]
            c_n0 = O.news_graph_context(P, a[0], a[2])                      # util.py:37-44 (the cached initial context)
            cn, cu = O.inference(P, *a, c_n0)
            out.append(O.logits(cn, cu))
            ctx.append((cn * cu).abs().sum(1))
    return torch.cat(out).numpy(), torch.cat(ctx).numpy()


@pytest.mark.parametrize('workload', ['mind_small_dev_n3_L3', 'mind_small_dev_n5_L3', 'wide_n8_L7'])
def test_bench_batch_sampled_against_oracle(workload):
    import bench
    from digat_b200 import scoring
    from digat_b200.graphEncoders import DIGAT
    BATCH, STEPS, SAMPLE = 4096, 2, 256
    cfg, sd, corpus = bench.build_workload(workload)
    enc = DIGAT(cfg, 400)
    enc.load_state_dict(sd)
    enc = enc.cuda().eval()
    scorer = scoring.Scorer(enc, corpus, 'cuda:0')
    scorer.cache_news_context()
    beh = torch.from_numpy(corpus.pair_behavior[:BATCH * STEPS]).cuda()
    news = torch.from_numpy(corpus.pair_news[:BATCH * STEPS]).cuda()
    res = torch.cat(scoring.score_resident_batches(
        scorer, ((beh[s * BATCH:(s + 1) * BATCH], news[s * BATCH:(s + 1) * BATCH]) for s in range(STEPS))))
    host = [scoring.host_batch(corpus, np.arange(s * BATCH, (s + 1) * BATCH), pin=True) for s in range(STEPS)]
    hst = torch.cat(scoring.score_host_batches(scorer, host))
    torch.cuda.synchronize()
    scorer.check_index_errors()
    assert res.shape == (BATCH * STEPS,) and bool(torch.isfinite(res).all())
    # both drivers run the same rows through the same kernels
    assert torch.equal(res, hst)
    rng = np.random.Generator(np.random.PCG64(17))
    ids = np.sort(rng.choice(BATCH * STEPS, size=SAMPLE, replace=False))
    ref32, terms = _oracle_logits(sd, corpus, ids, torch.float32)
    ref64, terms64 = _oracle_logits(sd, corpus, ids, torch.float64)
    got = res.cpu().numpy()[ids]
    s32, s64, sref = err_stats(got, ref32, terms64), err_stats(got, ref64, terms64), err_stats(ref32, ref64, terms64)
    print('\n%s, %d of %d pairs, logits vs oracle fp32: %s\n    vs fp64: %s\n    oracle fp32 vs fp64: %s'
          % (workload, SAMPLE, BATCH * STEPS, fmt_stats(s32), fmt_stats(s64), fmt_stats(sref)))
    assert s32['max_norm'] < LOGIT_TOL, s32
    assert s64['max_norm'] < LOGIT_TOL, s64
    assert s32['cond'] < COND_TOL and s64['cond'] < COND_TOL, (s32, s64)
    # ranking agreement on the sampled impressions (north star: AUC/MRR/nDCG equal to 1e-4): the order of the sampled
    # pairs inside each impression is the oracle's wherever the oracle's own fp32-vs-fp64 uncertainty does not exceed the gap
    b = corpus.pair_behavior[ids]
    for imp in np.unique(b):
        sel = np.nonzero(b == imp)[0]
        if len(sel) < 2:
            continue
        o_ref, o_got = np.argsort(-ref64[sel], kind='stable'), np.argsort(-got[sel], kind='stable')
        gaps = np.abs(np.diff(ref64[sel][o_ref]))
        if gaps.size and gaps.min() > 1e-4 * max(np.abs(ref64[sel]).max(), 1e-30):
            assert np.array_equal(o_ref, o_got)
