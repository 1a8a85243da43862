"""Development timing probe (not the contract bench): the half-list kernels against each other on one system."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _pkg; _pkg.load()
from dl_poly_b200 import engine, systems

name = sys.argv[1] if len(sys.argv) > 1 else "ionic_1m"
s = systems.by_name(name)
sr = engine.ShortRange(0)
sr.dev_setup_system(s)
sr.dev_load_atoms(s.xyz, s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
sr.dev_relocate_serial(); sr.dev_halo_serial()
for which in (0, 1, 2, 0, 1, 2):
    sr.set_list_kernel(which)
    ts, tk = [], []
    for rep in range(5):
        sr.dev_link_cell_pairs()
        t = sr.last_timings()
        ts.append(t["list_ms"]); tk.append(t["full_list_kernel_ms"])
    out = sr.dev_two_body_forces()
    print("%s list kernel %d: build ms min %.3f (list kernel %.3f)  pairs %d  engvdw %.12e engcpe %.12e" %
          (name, which, min(ts), min(tk), sr.dev_list_pairs(), out[0], out[2]), flush=True)
sr.close()
