"""Per-entry-point GPU time of one eager training step (bench.py --mode train shapes) + total GPU time of the step."""
import collections, sys
import numpy as np, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from digat_b200 import _lib, synth
from digat_b200.model import Model
D = 400
cfg = synth.make_config(SAG_neighbors=3, SAG_hops=2, graph_depth=3, dropout_rate=0.2)
sd = synth.make_state_dict(cfg, D=D, seed=0)
corpus = synth.make_corpus(cfg, D=D, n_news=20000, n_behaviors=4096, mean_candidates=8.0, seed=0)
dev = torch.device('cuda:0')
model = Model(cfg, D); model.graph_encoder.load_state_dict(sd); model = model.to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
rng = np.random.Generator(np.random.PCG64(0))
emb = torch.from_numpy(corpus.news_embeddings).to(dev); node = torch.from_numpy(corpus.news_node_ID.astype(np.int64)).to(dev)
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
    opt.zero_grad(set_to_none=True); loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0); opt.step()
    return loss
for _ in range(3): step(inputs())
torch.cuda.synchronize()
inp = inputs()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.start_profile()
e0.record(); step(inp); e1.record(); torch.cuda.synchronize()
rec = _lib.stop_profile()
agg = collections.defaultdict(lambda: [0, 0.0])
for name, a, ms in rec:
    key = name
    if name in ('digat_linear_f32', 'digat_linear_tf32x3'):
        M, N, K = (a[7], a[8], a[9]) if name == 'digat_linear_f32' else (a[8], a[9], a[10])
        key = '%s[M=%d,N=%d,K=%d]' % (name, M, N, K)
    if name == 'digat_linear_wgrad': key = '%s[M=%d,N=%d,K=%d]' % (name, a[6], a[7], a[8])
    if name in ('digat_graph_layer_fwd', 'digat_graph_layer_bwd'): key = '%s[n=%d]' % (name, a[7] if name.endswith('fwd') else a[13])
    agg[key][0] += 1; agg[key][1] += ms
tot = sum(v[1] for v in agg.values())
print('step wall (events, profiling on) %.2f ms; sum of our kernels %.2f ms over %d launches' % (e0.elapsed_time(e1), tot, len(rec)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print('%7.3f ms  x%3d  %s' % (v[1], v[0], k))
