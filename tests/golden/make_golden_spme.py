"""Generates tests/golden/spme_*.json from oracle/spme_oracle.py (run from the repo root: python tests/golden/make_golden_spme.py).

The fixtures freeze the numpy restatement's answers (pinned to the exact Ewald reciprocal sum by tests/test_spme_oracle.py), so a
change of the restatement shows; tests/test_spme_oracle.py::test_golden_spme_fixture reads them back."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import _pkg  # noqa: E402

_pkg.load()
from dl_poly_b200 import dd, systems  # noqa: E402
from oracle import spme_oracle as so  # noqa: E402

CASES = [("spme_nacl_512_order8", "nacl", dict(ncell=4, rcut=8.0, padding=0.2), 8),
         ("spme_water_1536_order6", "spce_water", dict(nmol=512, rcut=8.0, padding=0.2), 6)]
for name, gen, kw, nspl in CASES:
    s = getattr(systems, gen)(**kw)
    xyz = dd.read_config_fold(s.xyz, s.cell)[0]
    q = s.charge_site[s.lsite - 1]
    alpha, kdim = so.spme_grid(1.0e-6, s.rcut, s.cell)
    r = so.ewald_spme_forces_coul(s.cell, xyz, q, alpha, kdim, nspl, s.ff.scaling)
    g = dict(generator=gen, kwargs=kw, nspl=nspl, alpha=alpha, kdim=list(kdim), engcpe_rc=r["engcpe_rc"], vircpe_rc=r["vircpe_rc"],
             eng_recip=r["eng_recip"], stress=[float(v) for v in r["stress"]], force_l1=float(np.abs(r["forces"]).sum()),
             force_first=[float(v) for v in r["forces"][0]], force_last=[float(v) for v in r["forces"][-1]])
    json.dump(g, open(os.path.join(os.path.dirname(__file__), name + ".json"), "w"), indent=1)
    print(name, kdim, r["engcpe_rc"])
