"""Worker for tests/test_gpu_ddp.py (launched by torch.distributed.run, one rank per GPU, NCCL)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from digat_b200 import synth  # noqa: E402
from digat_b200.graphEncoders import DIGAT  # noqa: E402

ORDER = ('news_graph_embeddings', 'news_graph', 'news_graph_mask', 'user_news_embedding', 'user_graph',
         'user_category_mask', 'user_category_indices')


def main():
    out_path = sys.argv[1]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
    dev = torch.device('cuda', int(os.environ['LOCAL_RANK']))
    dist.init_process_group('nccl', device_id=dev)
    cfg = synth.make_config(graph_depth=2, dropout_rate=0.0)
    sd = synth.make_state_dict(cfg, seed=6)
    corpus = synth.make_corpus(cfg, n_news=300, n_behaviors=16, mean_candidates=3.0, seed=3)
    rows = 8 * world
    batch = synth.make_batch(corpus, np.arange(rows))
    m = DIGAT(cfg, 400)
    m.load_state_dict(sd)
    m = m.to(dev).train()
    ddp = DDP(m, device_ids=[dev.index])
    lo, hi = rank * 8, (rank + 1) * 8                      # DistributedSampler-style disjoint shards
    b = [batch[k][lo:hi].to(dev) for k in ORDER]
    cn, cu = ddp(*b)
    loss = (cn * cu).sum(1).mean()
    loss.backward()                                        # DDP all-reduces (averages) the gradients over NCCL
    torch.cuda.synchronize()
    if rank == 0:
        g_ddp = {k: v.grad.detach().cpu() for k, v in m.named_parameters()}
        # single-process reference on the union batch, same kernels
        m1 = DIGAT(cfg, 400)
        m1.load_state_dict(sd)
        m1 = m1.to(dev).train()
        bb = [batch[k].to(dev) for k in ORDER]
        cn, cu = m1(*bb)
        (cn * cu).sum(1).mean().backward()
        torch.cuda.synchronize()
        worst = 0.0
        for k, v in m1.named_parameters():
            a, r = g_ddp[k].double(), v.grad.detach().cpu().double()
            worst = max(worst, float((a - r).abs().max() / r.abs().max().clamp_min(1e-30)))
        with open(out_path, 'w') as f:
            f.write('%.6e %d\n' % (worst, world))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
