"""Development timing probe (not the contract bench): kernel times for one system."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _pkg; _pkg.load()
from dl_poly_b200 import engine, systems

name = sys.argv[1] if len(sys.argv) > 1 else "ionic_1m"
modes = [int(m) for m in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1"])]
s = systems.by_name(name)
print(name, s.megatm, "atoms  max_list", s.max_list, flush=True)
for mode in modes:
    sr = engine.ShortRange(0)
    if mode == modes[0]:
        print("fp64 peak TFLOP/s:", sr.fp64_peak(0.3))
    sr.dev_setup_system(s)
    sr.set_force_mode(mode)
    sr.dev_load_atoms(s.xyz, s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
    sr.dev_relocate_serial(); sr.dev_halo_serial()
    print(" mode", mode, "counts", sr.dev_counts())
    for rep in range(3):
        sr.dev_link_cell_pairs()
        t = sr.last_timings()
        print("  list build ms %.3f (kernel %.3f)" % (t["list_ms"], t["full_list_kernel_ms"]))
    for rep in range(5):
        out = sr.dev_two_body_forces()
        t = sr.last_timings()
        print("  forces ms %.3f (pair kernel %.3f)  -> %.3f G atom-steps/s (force only)" % (t["force_ms"], t["pair_kernel_ms"], s.megatm / t["force_ms"] / 1e6))
    print("  out", out[:6])
    sr.close()
