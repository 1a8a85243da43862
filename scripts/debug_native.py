import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import _pkg; _pkg.load()
from dl_poly_b200 import engine, systems
from oracle import oracle as ora
from util import world_for
for name, s in [("argon", systems.argon(6)), ("nacl", systems.nacl(4, rcut=8.0, padding=0.2)), ("water", systems.spce_water(512, rcut=8.0, padding=0.2))]:
    for rep in range(3):
        w = world_for(s, P=1, with_halo=False, with_list=False)
        w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
        sr = engine.ShortRange(0)
        sr.dev_setup_system(s)
        sr.dev_load_atoms(s.xyz, s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
        sr.dev_relocate_serial(); sr.dev_halo_serial()
        po, pg = w.parts(0), sr.dev_get_parts()
        io, ig = w.ints(0), sr.dev_get_ints()
        print(name, rep, "counts", sr.dev_counts(), w.counts(0)["natms"], w.counts(0)["nlast"])
        for k in ("xxx", "yyy", "zzz", "chge"):
            bad = np.nonzero(po[k] != pg[k])[0]
            if len(bad):
                print("  ", k, "mismatch", len(bad), "first", bad[:5], po[k][bad[:3]], pg[k][bad[:3]], (po[k][bad[:3]] - pg[k][bad[:3]]), "ltg", io["ltg"][bad[:3]], ig["ltg"][bad[:3]], "natms", s.megatm)
        for k in ("ltg", "ixyz"):
            bad = np.nonzero(io[k] != ig[k])[0]
            if len(bad): print("  ", k, "mismatch", len(bad), bad[:5])
        sr.close()
# c3 list
s = systems.by_name("c3")
w = world_for(s, P=1, with_halo=False, with_list=False)
w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
sr = engine.ShortRange(0); sr.dev_setup_system(s)
sr.dev_load_atoms(s.xyz, s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs(want_ref_list=True)
ref, got = w.list(0), sr.dev_get_list()
bad = np.nonzero((ref[:, :4] != got[:, :4]).any(1))[0]
print("c3 rows with different counters:", len(bad), bad[:10])
for i in bad[:3]:
    print(i, ref[i, :4], got[i, :4])
    a, b = set(ref[i, 4:4 + ref[i, 1]].tolist()), set(got[i, 4:4 + got[i, 1]].tolist())
    print("  only ref", sorted(a - b)[:10], "only got", sorted(b - a)[:10])
po, pg = w.parts(0), sr.dev_get_parts()
print("c3 pos mismatch", [(k, int((po[k] != pg[k]).sum())) for k in ("xxx", "yyy", "zzz")])
# empty domain
from dl_poly_b200.lib import COREPART
s = systems.argon(6)
sr = engine.ShortRange(0); sr.set_cell(s.cell, s.imcon); sr.set_cutoffs(s.rcut, s.padding, s.pdplnc); sr.set_forcefield(s.ff)
parts = np.zeros(0, dtype=COREPART)
try:
    lst = sr.link_cell_pairs(0, 0, parts, np.zeros(0, np.int32), np.zeros(0, np.int32), max_list=s.max_list)
    print("empty ok", lst.shape, sr.two_body_forces(0, 0, parts))
except Exception as e:
    print("empty domain error:", e)
