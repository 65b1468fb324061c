"""N > 1 on real GPUs (skipped on a 1-GPU box): partitioned train step vs the oracle, via torchrun — the 1-D row
partition and, with GNN_GRID, the 2-D partition of csrc/trainer_grid.cu (tests/dist_check.py does the checking)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,grid", [(2, "row"), (2, "1x2"), (2, "2x1"), (4, "2x2"), (8, "2x4")])
def test_partitioned_train_step_matches_oracle(world, grid):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, GNN_GRID=grid))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
