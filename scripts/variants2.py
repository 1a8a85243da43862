"""Development probe (results of the variants are not meaningful): pair-kernel time on the fp32-h layout against the L1
size (DLPGPU_SMEM_PAD takes shared memory away from the L1) and with a conflict-free Ewald table index (DLPGPU_VARIANT=128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _pkg; _pkg.load()
from dl_poly_b200 import engine, systems

s = systems.by_name(sys.argv[1] if len(sys.argv) > 1 else "ionic_1m")
cases = [(v, p) for p in (0, 60, 135) for v in (0x801, 128)]
for v, pad in cases:
    os.environ["DLPGPU_VARIANT"] = str(v)
    os.environ["DLPGPU_SMEM_PAD"] = str(pad)
    sr = engine.ShortRange(0)
    sr.dev_setup_system(s); sr.set_force_mode(1)
    sr.dev_load_atoms(s.xyz, s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
    sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
    ts = []
    for rep in range(5):
        sr.dev_two_body_forces(); ts.append(sr.last_timings()["pair_kernel_ms"])
    print("variant %5d  smem pad %3d KB  pair kernel ms: min %.4f  %s" % (v, pad, min(ts), ["%.3f" % t for t in ts]), flush=True)
    sr.close()
