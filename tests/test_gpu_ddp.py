"""Multi-GPU training path (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests -m gpu`): gradients after
DistributedDataParallel's NCCL all-reduce over k ranks equal the single-process gradients on the union batch
(reference trainer.py:19 wraps the model in DDP)."""
import os
import subprocess
import sys

import pytest
import torch

from tests.helpers import ROOT

pytestmark = pytest.mark.gpu


def test_ddp_gradients_match_single_process(tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs (run under gpurun --gpus 2)')
    world = 2
    out = tmp_path / 'ddp.txt'
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % world,
                        '--master-addr', '127.0.0.1', '--master-port', '29541',
                        os.path.join(ROOT, 'tests', 'ddp_worker.py'), str(out)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    worst, w = out.read_text().split()
    assert int(w) == world
    assert float(worst) < 1e-4, 'DDP gradients differ from single-process gradients: rel %.3e' % float(worst)
