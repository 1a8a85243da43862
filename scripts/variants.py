"""Development probe: pair-kernel time under the DLPGPU_VARIANT timing experiments (results of variants are garbage)."""
import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import _pkg; _pkg.load()
    from dl_poly_b200 import engine, systems
    s = systems.by_name(sys.argv[2])
    sr = engine.ShortRange(0)
    sr.dev_setup_system(s); sr.set_force_mode(1)
    sr.dev_load_atoms(s.xyz, s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
    sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
    ts = []
    for rep in range(6):
        sr.dev_two_body_forces(); ts.append(sr.last_timings()["pair_kernel_ms"])
    print("variant %s pair kernel ms: min %.4f  all %s" % (os.environ.get("DLPGPU_VARIANT", "0"), min(ts), ["%.3f" % t for t in ts]), flush=True)
    sr.close()
else:
    name = sys.argv[1] if len(sys.argv) > 1 else "ionic_1m"
    for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0", "1", "2", "4", "8", "3", "5", "7", "13"]):
        env = dict(os.environ, DLPGPU_VARIANT=v)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", name], env=env)
