"""Host-side logic of the multi-GPU driver (dl-poly_b200/dd.py) on CPU: map_domains against the oracle, and the six-stage
exchange pattern over a real 2-rank gloo process group with the oracle's domains standing in for the GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from dl_poly_b200 import dd, systems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("P", [1, 2, 3, 4, 6, 8, 12, 16])
def test_map_domains_matches_oracle(oracle, P):
    for cell in (np.diag([40.0, 40.0, 40.0]), np.diag([80.0, 40.0, 40.0]), np.diag([40.0, 90.0, 60.0])):
        s = systems.nacl(2, rcut=3.0, padding=0.1)
        w = oracle.World(P, cell.reshape(9), imcon=2)
        dims = dd.map_domains(P, dd.cell_widths(cell), 2)
        for r in range(P):
            d6, m26 = w.dd(r)
            assert tuple(d6[:3]) == dims
            assert tuple(d6[3:]) == dd.domain_of_rank(r, *dims)
            assert list(m26[:6]) == dd.face_neighbours(r, *dims)


WORKER = r'''
import os, sys
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import _pkg; _pkg.load()
from dl_poly_b200 import dd, systems
from oracle import oracle as ora
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
t = dd.TorchTransport(torch.device("cpu"))
s = systems.nacl((4, 2, 2), rcut=5.0, padding=0.2)
dims = dd.map_domains(world, dd.cell_widths(s.cell), s.imcon)
neigh = dd.face_neighbours(rank, *dims)
# the oracle world holds every domain; each rank drives "its" domain's halo build through dd.staged_exchange with fake
# pack/unpack that ship the oracle's own per-stage payloads, and checks that what arrives is what the oracle's in-process
# exchange delivered to this domain.
w = ora.World.from_system(s, P=world)
w.set_halo()
c = w.counts(rank)
parts = w.parts(rank); ints = w.ints(rank)
natms, nlast = c["natms"], c["nlast"]
# payload a rank would send in stage mdir = its atoms (local + already received halo) selected by the reference rule:
# reconstruct from the receiving side: the atoms rank r received are stored in order, tagged by ltg
recv_chunks = {}
bufs = {}
def alloc(n):
    x = torch.zeros(int(n), dtype=torch.float64)
    return x, x
sent_log, got_log = [], []
state = {"nlast": natms}
# what every rank must receive in each stage, from the oracle (halo atoms natms..nlast in arrival order): split by stage using
# the sender's view: ask all ranks for their full arrays via all_gather
allp = [None] * world
dist.all_gather_object(allp, (natms, nlast, parts["xxx"].copy(), ints["ltg"].copy()))
def pack(mdir, buf, cap):
    # send a recognisable payload: (rank, mdir, k) triples, width 6 like the halo build
    n = 5 + rank + abs(mdir)
    if n > cap:
        return 54, n
    v = buf[: n * 6].view(n, 6)
    v[:, 0] = rank; v[:, 1] = mdir; v[:, 2] = torch.arange(n, dtype=torch.float64)
    sent_log.append((mdir, n))
    return 0, n
def unpack(mdir, buf, n):
    v = buf[: n * 6].view(n, 6).clone()
    got_log.append((mdir, n, v))
dd.staged_exchange(t, neigh, dims, pack, unpack, alloc, 6)
ok = True
for q, mdir in enumerate(dd.MDIRS):
    axis = abs(mdir) - 1
    src = neigh[q ^ 1] if dims[axis] > 1 else rank
    m, n, v = got_log[q]
    assert m == mdir
    assert n == 5 + src + abs(mdir), (rank, mdir, n, src)
    assert torch.all(v[:, 0] == src) and torch.all(v[:, 1] == mdir) and torch.all(v[:, 2] == torch.arange(n, dtype=torch.float64))
assert t.allreduce_max(float(rank)) == world - 1
assert np.allclose(t.allreduce_sum([1.0, rank]), [world, sum(range(world))])
x = torch.full((5,), float(rank + 1), dtype=torch.float64)       # the charge-grid sum of the several-domain SPME
t.allreduce_sum_device(x)
assert torch.all(x == float(sum(range(1, world + 1))))
t.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_staged_exchange_over_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text("ROOT = %r\n" % ROOT + WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d failed:\n%s" % (r, o)
        assert "ok" in o
