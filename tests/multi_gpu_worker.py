"""Worker of tests/test_gpu_multi.py: one rank per GPU under torch.distributed.run.  Every rank runs the sharded
vt_fit and the sharded rotation x translation search with the LIBRARY's NCCL communicator and writes what it got."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def inputs():
    from oracle.pyoracle import Oracle
    from tests.synth import make_structure, make_target
    orc = Oracle()
    pos, radii, mass = make_structure(1500, 14.0, seed=41)
    g, t = make_target(orc, pos, radii, 5.0, 1.0, rot=(48, 24, 312), noise=0.02, seed=42, mass=mass)
    return pos, radii, mass, g, t


def run(vt, pos, radii, mass, g, t):
    target = vt.VolMap(vt.Grid.from_fields(g), t)
    newpos, res = vt.fit(pos, radii, mass, target, 5.0)
    com = vt.measure_center(pos, mass)
    ctx = vt.FitContext(pos, radii, target, 5.0)
    rots = np.array([[0, 0, 0], [48, 24, 312], [24, 48, 0], [48, 24, 300], [40, 20, 310]], np.float32)
    bcc, bidx, bpose, _ = ctx.search_grid(com, rots, 1, 1.0)
    ctx.close()
    return {"pos_bits": newpos.view(np.uint32).tolist()[:3000], "pos_sum": float(newpos.astype(np.float64).sum()),
            "best_cc": float(res.best_cc), "stage_best": list(res.stage_best), "n_evaluated": int(res.n_evaluated),
            "n_computed": int(res.n_computed), "grid_cc": float(bcc), "grid_idx": int(bidx), "grid_pose": bpose.tolist()}


def main():
    import torch
    import torch.distributed as dist
    import voltool_b200 as vt
    from voltool_b200 import dist as vdist
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    vt.init(local)
    rank, world = vdist.init_nccl_from_torch()
    out = run(vt, *inputs())
    out["rank"], out["world"] = rank, world
    with open(os.path.join(sys.argv[1], f"rank{rank}.json"), "w") as f:
        json.dump(out, f)
    vt.nccl_shutdown()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
