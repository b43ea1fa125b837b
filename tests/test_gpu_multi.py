"""The sharded pose search on real GPUs (skips below 2): N ranks under torch.distributed.run, the library's own NCCL
communicator (vt_nccl_init), vt_fit's five rotate() tables merged by one ncclAllGather each and the rotation x
translation search by one more -- every rank must return exactly what a single rank returns, bit for bit."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_fit_equals_single_rank(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    import voltool_b200 as vt
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests.multi_gpu_worker import inputs, run
    vt.init(0)
    single = run(vt, *inputs())
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                           "--master-addr", "127.0.0.1", "--master-port", "29533",
                           os.path.join(ROOT, "tests", "multi_gpu_worker.py"), str(tmp_path)], env=env, timeout=600)
    for r in range(world):
        got = json.load(open(tmp_path / f"rank{r}.json"))
        assert got["world"] == world
        for key in ("pos_bits", "best_cc", "stage_best", "n_evaluated", "n_computed", "grid_cc", "grid_idx", "grid_pose"):
            assert got[key] == single[key], (r, key)
