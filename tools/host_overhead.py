"""How much HOST time does one scoring step cost (enqueue only)?  Times the three host phases of the pipelined resident
driver -- begin_resident (flag kernels), finish_prepare minus its event wait, score_prepared (encoder launches) -- per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from digat_b200 import scoring
from digat_b200.graphEncoders import DIGAT

cfg, sd, corpus = bench.build_workload('mind_small_dev_n3_L3')
enc = DIGAT(cfg, 400); enc.load_state_dict(sd); enc = enc.cuda().eval()
scorer = scoring.Scorer(enc, corpus, 'cuda:0'); scorer.cache_news_context()
B = 4096
beh = torch.from_numpy(corpus.pair_behavior[:B * 30]).cuda(); news = torch.from_numpy(corpus.pair_news[:B * 30]).cuda()
t_begin, t_fin, t_wait, t_score = [], [], [], []
orig_sync = torch.cuda.Event.synchronize
waits = [0.0]
def timed_sync(self):
    t = time.perf_counter(); orig_sync(self); waits[0] += time.perf_counter() - t
torch.cuda.Event.synchronize = timed_sync
for s in range(30):
    t0 = time.perf_counter()
    begun = scorer.begin_resident(beh[s * B:(s + 1) * B], news[s * B:(s + 1) * B])
    t1 = time.perf_counter()
    waits[0] = 0.0
    prep = scorer.finish_prepare(begun)
    t2 = time.perf_counter()
    out = scorer.score_prepared(prep)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    if s >= 5:
        t_begin.append(t1 - t0); t_fin.append(t2 - t1 - waits[0]); t_wait.append(waits[0]); t_score.append(t3 - t2)
ms = lambda v: 1e3 * float(np.median(v))
print('host ms per step (median of 25): begin_resident %.3f, finish_prepare (without its event wait) %.3f, score_prepared %.3f, total %.3f  (cores: %d)'
      % (ms(t_begin), ms(t_fin), ms(t_score), ms(t_begin) + ms(t_fin) + ms(t_score), os.cpu_count()))
