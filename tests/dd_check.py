"""Multi-GPU parity check (run under torch.distributed.run, one rank per GPU): see tests/dd_common.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist

import _pkg

_pkg.load()
from dl_poly_b200 import dd
import dd_common

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t = dd.TorchTransport(torch.device("cuda", local))
which = sys.argv[1] if len(sys.argv) > 1 else "nacl"
rep = dd_common.check_rank(t, local, which)
spme = None
if which == "nacl":     # the several-domain SPME (replicated grid, NCCL all-reduce of the charge grids) against the one-domain oracle
    mine = dd_common.spme_rank(t, local, which)
    reps = [None] * world
    dist.all_gather_object(reps, mine)
    if rank == 0:
        spme = dd_common.spme_compare(reps, which)
if rank == 0:
    print("dd_check %s world=%d %s spme=%s OK" % (which, world, rep, spme))
dist.destroy_process_group()
