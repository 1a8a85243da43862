"""Two thread-ranks on one GPU: domain creation + first rebuild with a short peer timeout (diagnostic)."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import _pkg; _pkg.load()
import torch
from dl_poly_b200 import dd
import dd_common

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
grp = dd.ThreadGroup(world)
s = dd_common.make_system("nacl")
def body(rank):
    try:
        torch.cuda.set_device(0)
        t = grp.transport(rank)
        dom = dd.Domain(s, device=0, transport=t)
        dom.sr.dev_xchg_set_timeout(float(os.environ.get("PROBE_TIMEOUT", "5")))
        t.barrier()
        t0 = time.time()
        dom.rebuild()
        print("rank", rank, "rebuild ok", dom.sr.dev_counts(), time.time() - t0, flush=True)
        out = dom.forces()
        for k in range(5):
            dom.step(0.002, lazy=True)
        print("rank", rank, "steps ok", dom.collect()[:3], flush=True)
    except BaseException as e:
        print("rank", rank, "FAILED", repr(e)[:900], flush=True)
        grp.bar.abort()
th = [threading.Thread(target=body, args=(r,)) for r in range(world)]
[x.start() for x in th]; [x.join() for x in th]
