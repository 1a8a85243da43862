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
