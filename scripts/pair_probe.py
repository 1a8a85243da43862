"""Development timing probe (not the contract bench): the pair kernels against each other on one system."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _pkg; _pkg.load()
from dl_poly_b200 import engine, systems

name = sys.argv[1] if len(sys.argv) > 1 else "ionic_1m"
whiches = [int(m) for m in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "2"])]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
s = systems.by_name(name)
sr = engine.ShortRange(0)
sr.dev_setup_system(s)
sr.dev_load_atoms(s.xyz, s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
sr.dev_relocate_serial(); sr.dev_halo_serial()
sr.dev_link_cell_pairs()
t = sr.last_timings()
print(name, s.megatm, "atoms; list build ms %.3f (kernel %.3f)" % (t["list_ms"], t["full_list_kernel_ms"]), flush=True)
ref = None
for which in whiches:
    sr.set_pair_kernel(which=which)
    ts = []
    for rep in range(reps):
        out = sr.dev_two_body_forces()
        ts.append(sr.last_timings()["pair_kernel_ms"])
    used, terr = sr.pair_kernel_used()
    print(" which %d -> kernel %d (packed table error %.2e): pair kernel ms min %.4f median %.4f" % (which, used, terr, min(ts), float(np.median(ts))))
    if ref is None:
        ref = out
    else:
        print("   sums vs first: max rel %.2e" % np.max(np.abs(out[:6] - ref[:6]) / np.maximum(np.abs(ref[:6]), 1e-300)))
sr.close()
