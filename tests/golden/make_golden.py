"""Generates tests/golden/*.json from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference itself cannot be built in this image (no Fortran compiler) and its regression inputs are downloaded at
build time, so these fixtures freeze the ORACLE's outputs (already pinned to test_vdw.F90 and a brute force)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from dl_poly_b200 import systems  # noqa: E402
from oracle import oracle as ora  # noqa: E402

CASES = [("argon_864_P1", "argon", dict(ncell=6), 1),
         ("nacl_4096_P8", "nacl", dict(ncell=8, rcut=8.0, padding=0.2), 8),
         ("water_1536_P1", "spce_water", dict(nmol=512, rcut=8.0, padding=0.2), 1),
         ("nacl_table_512_P1", "nacl", dict(ncell=4, rcut=8.0, padding=0.2, tabfile=True, force_shift=True), 1)]
ora.build()
for name, gen, kw, P in CASES:
    s = getattr(systems, gen)(**kw)
    w = ora.World.from_system(s, P=P)
    w.set_halo()
    assert w.link_cell_pairs() == 0
    out = w.two_body()
    f = w.gather_forces()
    g = dict(generator=gen, kwargs=kw, P=P, nlast=[w.counts(r)["nlast"] for r in range(P)],
             list_entries=int(sum(w.list(r)[:, 1].sum() for r in range(P))), out=[float(v) for v in out],
             force_l1=float(np.abs(f).sum()))
    json.dump(g, open(os.path.join(os.path.dirname(__file__), name + ".json"), "w"), indent=1)
    print(name, g["nlast"], g["list_entries"])

# ---- host logic of the path: vdw_lrc, the end of two_body_forces, the vnl_check decisions (test_host_cpp.py holds the C++ and
# Python hosts against the oracle; this file freezes the oracle's own answers)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_host_cpp as th  # noqa: E402

host = {}
s = systems.nacl(3, rcut=9.0, padding=0.2)
num_type = [float((s.type_site[s.lsite - 1] == t).sum()) for t in (1, 2)]
e, v = th._ora_lrc(ora, s.ff, num_type, [3.0, 5.0], 1, s.volume)
host["vdw_lrc_nacl_216_rc9_frozen_3_5"] = [e, v]
s = systems.argon(4)
e2, v2 = th._ora_lrc(ora, s.ff, [float(s.megatm)], [0.0], 1, s.volume)
host["vdw_lrc_argon_256"] = [e2, v2]
tot, st = th._ora_epilogue(ora, [1.0e5, -2.0e5, -3.0e6, 2.5e6, 4.0e3, -5.0e3, -7.0e5, 6.0e5], True, 2.0, 0.3, 1.0, 5000.0, e, v, 8, np.arange(9.0))
host["epilogue_totals"] = [float(x) for x in tot]
host["epilogue_stress"] = [float(x) for x in st]
tols = [0.05, 0.02, 0.13, 0.16, 0.01, 0.18, 0.2, 0.0, 0.09, 0.17]
rows, kodes = th._ora_vnl_trace(ora, False, 0, 8.5, 0.1, np.diag([114.39] * 3).reshape(9), [1, 1, 1], tols)
host["vnl_tols"] = tols
host["vnl_trace_nostrict_argon"] = [[float(x) for x in r] for r in rows]
host["vnl_kodes"] = kodes
json.dump(host, open(os.path.join(os.path.dirname(__file__), "host_logic.json"), "w"), indent=1)
print("host_logic", host["vdw_lrc_argon_256"], kodes)
