"""Multi-GPU parity check (launch with torchrun, one rank per GPU): a mixed list of exploration sets is partitioned
over the ranks (some sets split between ranks), every rank sweeps its share, the per-set bests are all-gathered over
NCCL and combined on the device; every rank must end with the oracle's per-set maxima and the oracle's selection.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import RTOL, make_case, oracle_sweep  # noqa: E402
from cbo_with_oop_b200.engine import SetProblem, SweepEngine  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
specs = [dict(seed=61, N=96, d=1, c=1, n=8, p=(300,)), dict(seed=62, N=140, d=2, c=2, n=10, p=(40, 40)),
         dict(seed=63, N=200, d=3, c=1, n=12, p=(20, 20, 20)), dict(seed=64, N=70, d=2, c=0, n=9, p=(60, 30)),
         dict(seed=65, N=90, d=1, c=2, n=10, p=(640,), causal=False), dict(seed=66, N=150, d=2, c=1, n=11, p=(50, 50))]
cases = [make_case(**s) for s in specs]
best = float(min(np.min(k["y_int"]) for k, _ in cases))
eng = SweepEngine([SetProblem(**k) for k, _ in cases], device=f"cuda:{local}", rank=rank, world_size=world)
out = eng.sweep(best, "min")
refs = [oracle_sweep(o, best, "min") for _, o in cases]
vals = np.array([r["val"] for r in refs])
np.testing.assert_allclose(out.set_values, vals, rtol=RTOL)
np.testing.assert_array_equal(out.set_indices, [r["idx"] for r in refs])
assert out.set == int(np.argmax(vals)) and out.index == refs[out.set]["idx"]
split = [s for s in range(len(cases)) if 0 < eng.slices[s][1] < cases[s][0]["grid"][0].size * 0 + np.prod([len(t) for t in cases[s][0]["grid"]])]
print(f"rank {rank}/{world}: slices {[sl for sl in eng.slices]} split sets {split} -> set {out.set} index {out.index} OK", flush=True)
# every rank holds the same answer
t = torch.tensor([float(out.set), float(out.index), out.value], dtype=torch.float64, device=f"cuda:{local}")
g = [torch.empty_like(t) for _ in range(world)]
dist.all_gather(g, t)
assert all(torch.equal(g[0], x) for x in g)
dist.barrier()
dist.destroy_process_group()
