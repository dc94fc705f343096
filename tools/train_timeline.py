"""Kernel timeline of one replay of the graphed training step (torch.profiler / CUPTI): per-stream busy time, the gaps on
each stream and the largest ones.  python tools/train_timeline.py [out.csv]"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from digat_b200 import synth
    from digat_b200.model import Model
    from digat_b200.training import FlatAdam, GraphedTrainStep
    dev = torch.device('cuda:0')
    cfg = synth.make_config(SAG_neighbors=3, SAG_hops=2, graph_depth=3, dropout_rate=0.2)
    sd = synth.make_state_dict(cfg, D=400, seed=0)
    corpus = synth.make_corpus(cfg, D=400, n_news=20000, n_behaviors=4096, mean_candidates=8.0, seed=0)
    model = Model(cfg, 400)
    model.graph_encoder.load_state_dict(sd)
    model = model.to(dev).train()
    flat = FlatAdam(model.parameters(), lr=1e-4, max_norm=1.0)
    rng = np.random.Generator(np.random.PCG64(0))
    emb = torch.from_numpy(corpus.news_embeddings).to(dev)
    node = torch.from_numpy(corpus.news_node_ID.astype(np.int64)).to(dev)
    ng, nm = torch.from_numpy(corpus.news_graph).to(dev), torch.from_numpy(corpus.news_graph_mask).to(dev)
    hist = torch.from_numpy(corpus.history.astype(np.int64)).to(dev)
    ug, cm, ci = (torch.from_numpy(x).to(dev) for x in (corpus.user_graph, corpus.user_category_mask, corpus.user_category_indices))

    def inputs():
        beh = torch.from_numpy(rng.integers(0, hist.shape[0], size=64)).to(dev)
        cand = torch.from_numpy(rng.integers(1, emb.shape[0], size=(64, 5))).to(dev)
        return (emb[hist[beh]], ug[beh], cm[beh], ci[beh], emb[node[cand]], ng[cand], nm[cand])

    def step(*inp):
        logits = model.forward_embeddings(*inp)
        loss = (-F.log_softmax(logits, dim=1).select(1, 0)).mean()
        flat.zero_grad()
        loss.backward()
        flat.step()
        return loss

    inp = inputs()
    g = GraphedTrainStep(step, inp, modules=(model,))
    for _ in range(3):
        g(*inp)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        g(*inp)
        torch.cuda.synchronize()
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
            ev.append((e.time_range.start, e.time_range.end, getattr(e, 'device_resource_id', -1), e.name))
    ev.sort()
    out = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/train_timeline.csv'
    with open(out, 'w') as f:
        f.write('start_us,end_us,stream,name\n')
        for s, e_, st, n in ev:
            f.write('%.3f,%.3f,%s,"%s"\n' % (s, e_, st, n[:100].replace('"', "'")))
    if not ev:
        print('no CUDA events captured')
        return
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    print('%d kernel events, span %.1f us' % (len(ev), t1 - t0))
    streams = sorted(set(e[2] for e in ev))
    for st in streams:
        es = [e for e in ev if e[2] == st]
        busy = sum(e[1] - e[0] for e in es)
        print('stream %s: %d kernels, busy %.1f us, first %.1f last %.1f' % (st, len(es), busy, es[0][0] - t0, es[-1][1] - t0))
    # union busy time over all streams, and idle gaps of the whole GPU
    merged, cur_s, cur_e = [], ev[0][0], ev[0][1]
    for s, e_, _, _ in ev[1:]:
        if s > cur_e:
            merged.append((cur_s, cur_e)); cur_s, cur_e = s, e_
        else:
            cur_e = max(cur_e, e_)
    merged.append((cur_s, cur_e))
    busy_any = sum(e_ - s for s, e_ in merged)
    print('GPU busy (any stream) %.1f us, idle %.1f us in %d gaps' % (busy_any, (t1 - t0) - busy_any, len(merged) - 1))
    gaps = sorted(((merged[i + 1][0] - merged[i][1], merged[i][1] - t0) for i in range(len(merged) - 1)), reverse=True)
    print('largest idle gaps (us @ offset):', ', '.join('%.1f@%.0f' % g_ for g_ in gaps[:12]))
    hist_g = np.array([g_[0] for g_ in gaps])
    if len(hist_g):
        print('gap median %.2f us, mean %.2f us, sum of gaps < 5us: %.1f us' % (np.median(hist_g), hist_g.mean(), hist_g[hist_g < 5].sum()))


if __name__ == '__main__':
    main()
