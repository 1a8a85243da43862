"""Development timing probe (not the contract bench): the SPME reciprocal-space call next to the short-range call on one system."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import _pkg; _pkg.load()
from dl_poly_b200 import engine, systems, tables

name = sys.argv[1] if len(sys.argv) > 1 else "ionic_1m"
nspl = int(sys.argv[2]) if len(sys.argv) > 2 else 8
s = systems.by_name(name)
alpha, kdim = tables.spme_grid(1.0e-6, s.rcut, s.cell)
sr = engine.ShortRange(0)
sr.dev_setup_system(s)
sr.dev_load_atoms(s.xyz, s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
sr.set_spme(kdim, nspl)
out = sr.dev_spme_forces(s.megatm)
ts = []
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = sr.dev_spme_forces(s.megatm)
    torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
f = sr.dev_two_body_forces()
t = sr.last_timings()
print("%s: %d atoms, grid %s order %d: spme call ms min %.3f median %.3f (wall, synchronous); engcpe_rc %.10e vircpe_rc %.10e; "
      "short-range force call %.3f ms (pair kernel %.3f)" % (name, s.megatm, kdim, nspl, min(ts), float(np.median(ts)), out[0], out[1],
                                                              t["force_ms"], t["pair_kernel_ms"]))
sr.close()
