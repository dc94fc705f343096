"""One eager training step between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/train_launches.py
and `python tools/train_launches.py --summarise X` -> per-kernel totals (every kernel of the step, torch's included)."""
import collections
import csv
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def summarise(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = collections.defaultdict(lambda: [0.0, 0])
    for r in rows[1:]:
        if r[ix['Metric Name']] != 'gpu__time_duration.sum':
            continue
        v = float(r[ix['Metric Value']].replace(',', ''))
        unit = r[ix['Metric Unit']]
        us = v / 1000 if unit.startswith('ns') else (v if unit.startswith('us') else v * 1000)
        name = re.sub(r'\(.*', '', r[ix['Kernel Name']])
        name = re.sub(r'<.*', '', name)
        tot[name][0] += us
        tot[name][1] += 1
    total = sum(v[0] for v in tot.values())
    print('%d launches, %.1f us of kernel time (serialised, cold cache)' % (sum(v[1] for v in tot.values()), total))
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print('%9.1f us %5.1f%% %4d x  %s' % (v[0], 100 * v[0] / total, v[1], k[:110]))


def main():
    import numpy as np
    import torch
    import torch.nn.functional as F
    from digat_b200 import synth
    from digat_b200.model import Model
    from digat_b200.training import FlatAdam
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

    def step(inp):
        logits = model.forward_embeddings(*inp)
        loss = (-F.log_softmax(logits, dim=1).select(1, 0)).mean()
        flat.zero_grad()
        loss.backward()
        flat.step()
        return loss

    for _ in range(3):
        inp = inputs()
        step(inp)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    step(inp)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == '__main__':
    if len(sys.argv) > 2 and sys.argv[1] == '--summarise':
        summarise(sys.argv[2])
    else:
        main()
