"""CPU tests of host-side logic: ranking/metrics, sharding (world_size 2 over gloo), bench helpers."""
import os
import subprocess
import sys

import numpy as np
import torch

from oracle import digat_oracle as O
from tests.helpers import GOLDEN, ROOT


def test_rank_lists_and_metrics_match_oracle_and_golden():
    from digat_b200 import evaluate
    z = np.load(os.path.join(GOLDEN, 'metrics.npz'))
    ranks = evaluate.rank_lists(z['scores'], z['imp'])
    assert ranks == O.rank_lists(z['scores'], z['imp'])
    labels = [z['labels'][z['imp'] == i].tolist() for i in range(int(z['imp'][-1]) + 1)]
    m = evaluate.metrics(ranks, labels)
    assert np.allclose(m, z['metrics'], atol=1e-12), (m, z['metrics'])


def test_rank_lists_ties_keep_candidate_order():
    from digat_b200 import evaluate
    r = evaluate.rank_lists(np.array([0.5, 0.5, 0.9, 0.5], dtype=np.float32), np.array([0, 0, 0, 0]))
    assert r == [[2, 3, 1, 4]]
    assert r == O.rank_lists([0.5, 0.5, 0.9, 0.5], [0, 0, 0, 0])


def test_shard_range_partitions_in_order():
    from digat_b200.scoring import shard_range
    for n in (0, 1, 7, 100, 2717103):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]


_GLOO_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DIGAT_ROOT"])
from digat_b200.scoring import shard_range
from digat_b200 import evaluate
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.Generator(np.random.PCG64(0))
n = 1001
scores_all = rng.normal(size=n).astype(np.float32)          # stands in for the per-pair logits
imp = np.sort(rng.integers(0, 50, size=n))
lo, hi = shard_range(n, rank, world)
mine = torch.from_numpy(scores_all[lo:hi].copy())           # each rank "scores" only its contiguous shard
per = (n + world - 1) // world
buf = torch.zeros(per); buf[:hi - lo] = mine
gathered = [torch.zeros(per) for _ in range(world)]
dist.all_gather(gathered, buf)                              # result collection only -- not on the data path
full = torch.cat([g[:shard_range(n, r, world)[1] - shard_range(n, r, world)[0]] for r, g in enumerate(gathered)]).numpy()
assert np.array_equal(full, scores_all), "sharded scores do not concatenate back in order"
assert evaluate.rank_lists(full, imp) == evaluate.rank_lists(scores_all, imp)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_pair_sharding_world_size_2_gloo(tmp_path):
    """The N>1 inference path: contiguous pair shards, no data-path collective, ordered concatenation."""
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, DIGAT_ROOT=ROOT)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29531', str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count('ok') == 2


def test_bench_algorithmic_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    # 3.77 / 4.57 / 14.95 MB per pair
    assert abs(bench.algorithmic_bytes_per_pair(10, 68, 3) / 1e6 - 3.77) < 0.02
    assert abs(bench.algorithmic_bytes_per_pair(26, 68, 3) / 1e6 - 4.57) < 0.02
    assert abs(bench.algorithmic_bytes_per_pair(65, 68, 7) / 1e6 - 14.95) < 0.05


def test_bench_reference_arm_runs_on_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '0'], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['value'] > 0 and line['cpu_baseline']['kind'] == 'port'


def test_similar_to_csr_and_impression_offsets():
    from digat_b200 import evaluate, graphs
    similar = [[], [(3, 0.9), (2, 0.5)], None, [(1, 0.75)]]
    off, idx, cos = graphs.similar_to_csr(similar, 5)                  # news 4 has no entry at all
    assert off.tolist() == [0, 0, 2, 2, 3, 3] and off.dtype == np.int64
    assert idx.tolist() == [3, 2, 1] and idx.dtype == np.int32
    assert cos.tolist() == [0.9, 0.5, 0.75] and cos.dtype == np.float64
    imp = np.array([0, 0, 0, 1, 3, 3])                                # impression 2 is empty
    assert evaluate.impression_offsets(imp).tolist() == [0, 3, 4, 4, 6]
    assert evaluate.impression_offsets(np.array([], dtype=np.int64)).tolist() == [0]
    # rank_lists / metrics on the same layout (host versions; the device versions are checked against them on the GPU)
    ranks = evaluate.rank_lists(np.array([0.1, 0.3, 0.3, 1.0, -1.0, 2.0], dtype=np.float32), imp)
    assert ranks == [[3, 1, 2], [1], [], [2, 1]]


_GLOO_FLATGRAD_WORKER = r'''
import os, sys
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
sys.path.insert(0, os.environ["DIGAT_ROOT"])
from digat_b200.training import FlatGradients, broadcast_parameters

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
torch.manual_seed(100 + rank)                                # different initial weights per rank on purpose
make = lambda: torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.ReLU(), torch.nn.Linear(16, 3))
ours = make()
broadcast_parameters(ours)                                   # every rank now holds rank 0's weights
ref = make()
ref.load_state_dict(ours.state_dict())
ddp = DDP(ref)
g = torch.Generator().manual_seed(7)
X = torch.randn(8 * world, 12, generator=g)
Y = torch.randn(8 * world, 3, generator=g)
xs, ys = X[rank * 8:(rank + 1) * 8], Y[rank * 8:(rank + 1) * 8]
flat = FlatGradients(ours.parameters())
for step in range(2):                                        # second pass: zero() + in-place accumulation into the views
    flat.zero()
    ((ours(xs) - ys) ** 2).mean().backward()
    flat.all_reduce_mean()
    ddp.zero_grad(set_to_none=True)
    ((ddp(xs) - ys) ** 2).mean().backward()
    for a, b in zip(ours.parameters(), ref.parameters()):
        assert a.grad.data_ptr() >= flat.flat.data_ptr() and a.grad.data_ptr() < flat.flat.data_ptr() + flat.flat.numel() * 4
        assert torch.allclose(a.grad, b.grad, rtol=1e-6, atol=1e-7), (step, (a.grad - b.grad).abs().max())
# and both equal the single-process gradient on the union batch
single = make()
single.load_state_dict(ours.state_dict())
((single(X) - Y) ** 2).mean().backward()
for a, b in zip(ours.parameters(), single.parameters()):
    assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-6)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_flat_gradient_all_reduce_equals_ddp_world_size_2_gloo(tmp_path):
    """The N>1 training path (digat_b200/training.py): one all-reduce of the flat gradient buffer gives DDP's gradients
    (reference trainer.py:19) and the single-process gradients of the union batch."""
    script = tmp_path / 'worker.py'
    script.write_text(_GLOO_FLATGRAD_WORKER)
    env = dict(os.environ, DIGAT_ROOT=ROOT)
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29533', str(script)],
                       env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count('ok') == 2
