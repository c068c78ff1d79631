"""torchrun check of the sharded selection path on real GPUs (one process per GPU, NCCL):
every rank scores its block of a common candidate set; the picks must equal the single-GPU picks on the whole set.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/check_sharded.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import pysot_b200 as pb
from pysot_b200.sharded import sharded_weighted_distance_merit

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for case, (d, n, m, box) in enumerate([(12, 300, 100003, "unit"), (30, 500, 50000, "iso"), (5, 64, 7001, "aniso")]):
    rng = np.random.default_rng(case)
    lb = np.zeros(d) if box == "unit" else (-15.0 * np.ones(d) if box == "iso" else rng.uniform(-5, 0, d))
    ub = np.ones(d) if box == "unit" else (20.0 * np.ones(d) if box == "iso" else lb + rng.uniform(1, 9, d))
    X = lb + (ub - lb) * rng.random((n, d))
    fX = np.sin(3 * ((X - lb) / (ub - lb)).sum(1))[:, None]
    Xpend = lb + (ub - lb) * rng.random((3, d))
    cand = lb + (ub - lb) * rng.random((m, d))
    cand[m - 5] = cand[7]  # duplicate across shards: the lower global index must win
    s = pb.GpuRBFInterpolant(dim=d, lb=lb, ub=ub, device=local)
    s.add_points(X, fX)
    w = [0.3, 0.5, 0.8, 0.95]
    bounds = np.linspace(0, m, world + 1).astype(int)
    bounds[1:-1] += 17 * np.arange(1, world)  # uneven shards
    mine = cand[bounds[rank]:bounds[rank + 1]]
    for dev_in in (False, True):
        inp = torch.from_numpy(mine).cuda() if dev_in else mine
        pts, idx = sharded_weighted_distance_merit(4, s, X, fX, inp, w, Xpend)
        ref_idx, _ = pb.merit_select_indices(4, s, X, cand, w, Xpend)
        good = idx.tolist() == ref_idx.tolist() and np.array_equal(pts, cand[ref_idx])
        ok &= good
        if rank == 0:
            print("case %d (d=%d n=%d m=%d %s, %s input, %d ranks): sharded %s single %s %s"
                  % (case, d, n, m, box, "device" if dev_in else "host", world, idx.tolist(), ref_idx.tolist(),
                     "OK" if good else "MISMATCH"), flush=True)
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ALL OK" if t.item() == 1.0 else "FAILED")
dist.destroy_process_group()
sys.exit(0 if t.item() == 1.0 else 1)
