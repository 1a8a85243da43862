import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import _pkg; _pkg.load()
from dl_poly_b200 import dd, engine, systems
name = sys.argv[1] if len(sys.argv) > 1 else "argon"
s = {"argon": lambda: systems.argon(6), "nacl": lambda: systems.nacl(4, rcut=8.0, padding=0.2), "water": lambda: systems.spce_water(512, rcut=8.0, padding=0.2)}[name]()
sr = engine.ShortRange(0)
sr.dev_setup_system(s)
sr.dev_load_atoms(dd.read_config_fold(s.xyz, s.cell)[0], s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
print(sr.dev_two_body_forces()[:6])
sr.close()
