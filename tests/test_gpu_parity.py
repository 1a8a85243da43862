"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): neighbour list bit-exact (we check the whole reference-format array, i.e. set, row
partition AND order); per-atom forces within 1e-9 relative; engsrp/engcpe/virial terms within 1e-10 relative.
"""
import numpy as np
import pytest

from dl_poly_b200 import dd, engine, systems
from dl_poly_b200.lib import COREPART
from util import domain_inputs, force_errors, parts_forces, per_atom_force_error, rel_err, world_for

pytestmark = pytest.mark.gpu

FORCE_TOL = 1.0e-9      # north star: per-atom forces, relative (to the largest force; the per-atom ratio is asserted looser below)
ENERGY_TOL = 1.0e-10    # north star: energies / virials, relative


def make_sr(s, dd=(1, 1, 1, 0, 0, 0)):
    sr = engine.ShortRange(0, dd)
    sr.set_cell(s.cell, s.imcon)
    sr.set_cutoffs(s.rcut, s.padding, s.pdplnc)
    sr.set_forcefield(s.ff)
    return sr


def check_dropin(s, P, ranks=None, mode=0):
    """Feeds each oracle domain through dlpgpu_link_cell_pairs / dlpgpu_two_body_forces exactly like the Fortran shim."""
    w = world_for(s, P=P)
    w.two_body()
    worst = dict(frel=0.0, fatom=0.0, erel=0.0)
    for r in (ranks if ranks is not None else range(P)):
        d = domain_inputs(w, r)
        sr = make_sr(s, d["dd"])
        sr.set_force_mode(mode)
        parts = d["parts"].copy()
        for k in ("fxx", "fyy", "fzz"):
            parts[k] = 0.0
        lst = sr.link_cell_pairs(d["natms"], d["nlast"], parts, d["ltype"], d["ltg"], d["lfrzn"], lbook=s.lbook, megfrz=s.megfrz,
                                 list_excl=d["list_excl"], max_list=d["max_list"])
        ref = w.list(r)
        assert lst.shape == ref.shape
        # bit-exact: counters (-3..0), members and order of the used part of every row
        assert np.array_equal(lst[:, :4], ref[:, :4])
        used = np.arange(ref.shape[1] - 4)[None, :] < ref[:, 1:2]
        assert np.array_equal(np.where(used, lst[:, 4:], 0), np.where(used, ref[:, 4:], 0))
        cells = sr.dev_get_cells()
        c = w.counts(r)
        assert (cells["nlx"], cells["nly"], cells["nlz"], cells["nlp"], cells["ncells"], cells["nsbcll"]) == \
            (c["nlx"], c["nly"], c["nlz"], c["nlp"], c["ncells"], c["nsbcll"])
        wc, al, ls = w.cells(r)
        assert np.array_equal(cells["which_cell"], wc)
        assert np.array_equal(cells["lct_start"] + 1, ls)                # 0-based slots vs Fortran 1-based
        assert np.array_equal(cells["at_list"] + 1, al)
        parts_again = parts.copy()
        # the call sequence of calculate_forces: the records link_cell_pairs uploaded are reused ...
        out = sr.two_body_forces(d["natms"], d["nlast"], parts, unchanged_since_list=True)
        # ... and the plain call (own upload) gives the same forces and sums up to the order of the atomic adds
        out_again = sr.two_body_forces(d["natms"], d["nlast"], parts_again)
        assert np.allclose(out_again, out, rtol=1e-12, atol=1e-12 * np.abs(out).max())
        fa, fb = parts_forces(parts_again, d["natms"]), parts_forces(parts, d["natms"])
        assert np.abs(fa - fb).max() <= 1e-12 * np.abs(fb).max()
        fo = parts_forces(d["parts"], d["natms"])
        fg = parts_forces(parts, d["natms"])
        a, b = force_errors(fg, fo)
        worst["frel"] = max(worst["frel"], a); worst["fatom"] = max(worst["fatom"], b)
        oo = w.results(r)
        scale = max(abs(oo[:6]).max(), 1.0)
        for k in range(6):
            e = abs(out[k] - oo[k]) / max(abs(oo[k]), 1e-6 * scale)
            worst["erel"] = max(worst["erel"], e)
            assert e <= ENERGY_TOL, ("energy/virial term", k, out[k], oo[k])
        sscale = abs(oo[6:15]).max()
        assert np.abs(out[6:15] - oo[6:15]).max() <= ENERGY_TOL * sscale
        assert a <= FORCE_TOL
        rep = per_atom_force_error(fg, fo)
        worst["fsig"] = max(worst.get("fsig", 0.0), rep["per_atom_significant"])
        # north star: per-atom relative 1e-9 for every atom whose net force is significant (>= 1e-3 of the largest); the ratio
        # over ALL atoms is bounded too (an atom whose pair terms cancel a million-fold keeps 1e-7)
        assert rep["per_atom_significant"] <= FORCE_TOL, rep
        assert b <= 1.0e-7, rep
        sr.close()
    return worst


@pytest.mark.parametrize("mode", [0, 1])
def test_argon_small_serial(mode):
    check_dropin(systems.argon(6), 1, mode=mode)


def test_argon_lj_direct_and_shifted():
    check_dropin(systems.argon(6, form="lj", direct=True, force_shift=True), 1)
    check_dropin(systems.argon(6, form="lj", force_shift=True), 1)


@pytest.mark.parametrize("mode", [0, 1])
def test_nacl_small_serial_and_domains(mode):
    check_dropin(systems.nacl(4, rcut=8.0, padding=0.2), 1, mode=mode)
    check_dropin(systems.nacl(8, rcut=8.0, padding=0.2), 8, mode=mode)


def test_nacl_table_file_force_shift():
    check_dropin(systems.nacl(4, rcut=8.0, padding=0.2, tabfile=True, force_shift=True), 1)


def test_nacl_vdw_cutoff_below_rcut_separate_grids():
    """rvdw < rcut: the vdW and Ewald tables have different grid spacings (the pair kernel's separate-grid instantiation)."""
    check_dropin(systems.nacl(4, rcut=8.0, padding=0.2, rvdw=6.5), 1, mode=1)


def test_nacl_ewald_only_and_partial_vdw():
    """No vdW potential at all (Ewald-only instantiation), and a force field where only Na-Cl carries one (entries with
    potential index 0 next to tabulated ones)."""
    check_dropin(systems.nacl(4, rcut=8.0, padding=0.2, vdw_pairs=()), 1, mode=1)
    check_dropin(systems.nacl(4, rcut=8.0, padding=0.2, vdw_pairs=((1, 2),)), 1, mode=1)


@pytest.mark.parametrize("kind,eps,damping", [("coul", 1.0, 0.0), ("dddp", 2.5, 0.0), ("fscp", 1.0, 0.0), ("fscp", 1.0, 0.2),
                                              ("rfp", 78.0, 0.0), ("rfp", 5.0, 0.25)])
def test_direct_space_coulomb_variants(kind, eps, damping):
    """coul_spole.F90: coul_cp / coul_dddp / coul_fscp / coul_rfp_forces (undamped and Fennell-Gezelter damped) instead of the
    Ewald term, with vdW tables next to them; SPC/E checks that excluded pairs simply drop out (no Ewald correction)."""
    check_dropin(systems.nacl(4, rcut=8.0, padding=0.2, coulomb=kind, eps=eps, damping=damping), 1, mode=1)
    check_dropin(systems.spce_water(512, rcut=8.0, padding=0.2, coulomb=kind, eps=eps, damping=damping), 1, mode=1)


def test_many_potentials_fit_the_fp32_h_layout():
    """Four species = ten vdW tables + Ewald: 424 KB in the fp64 table layout, 212 KB of g units in the fp32-h layout -- the
    fast kernel still applies; five species (16 tables) fall back to the general kernel with tables in global memory."""
    check_dropin(systems.ionic_mixture(4, ntypes=4), 1, mode=1)
    check_dropin(systems.ionic_mixture(3, ntypes=5), 1, mode=1)


def test_nacl_bhm_direct():
    check_dropin(systems.nacl(4, rcut=8.0, padding=0.2, direct=True), 1)


@pytest.mark.parametrize("mode", [0, 1])
def test_water_exclusions_subcells(mode):
    w = check_dropin(systems.spce_water(512, rcut=8.0, padding=0.2), 1, mode=mode)
    check_dropin(systems.spce_water(4096, rcut=8.0, padding=0.2), 8, ranks=[0, 7], mode=mode)


@pytest.mark.parametrize("pdplnc,nlp", [(2.0, 3), (0.5, 4)])
@pytest.mark.parametrize("mode", [0, 1])
def test_fine_subcells_nlp3_nir_shortcut(pdplnc, nlp, mode):
    """nlp >= 3 (neighbours.F90:537): cells within (nlp-1)^2 of the primary cell skip the distance test, so
    the reference list over-admits; the CUDA list must over-admit the SAME pairs (bit-exact list) and the forces must not see
    them.  nlp = 3 runs the warp-per-cell kernel's general loop, nlp = 4 the per-atom kernel (run table too long)."""
    s = systems.argon(6)
    s.pdplnc = pdplnc
    w = world_for(s, P=1)
    assert w.counts(0)["nlp"] == nlp
    check_dropin(s, 1, mode=mode)
    s = systems.nacl(4, rcut=8.0, padding=0.2)
    s.pdplnc = pdplnc / 2
    check_dropin(s, 1, mode=mode)


@pytest.mark.parametrize("mode", [0, 1])
def test_triclinic_cell_imcon3(mode):
    """Parallelepiped cell (imcon = 3): link cells in reduced coordinates (neighbours.F90:610-619), perpendicular widths from
    dcell, serial and 8 domains through the drop-in calls; native halo build + list + forces in the same cell."""
    s = systems.argon_triclinic(6)
    assert s.imcon == 3
    check_dropin(s, 1, mode=mode)
    check_dropin(systems.argon_triclinic(8), 8, ranks=[0, 3, 7], mode=mode)


def test_triclinic_native_trajectory():
    """imcon = 3 through the native path: halo images, pbcshift, vnl_check displacement with the parallelepiped minimum image,
    refresh and rebuild decisions track the oracle."""
    s = systems.argon_triclinic(6, temperature=300.0)
    dt = 0.004
    w = world_for(s, P=1, with_halo=False, with_list=False)
    sr = native_serial(s)
    w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0; w.two_body()
    sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs(); sr.dev_two_body_forces()
    po, pg = w.parts(0), sr.dev_get_parts()
    for k in ("xxx", "yyy", "zzz"):
        assert np.array_equal(po[k], pg[k]), k
    rebuilds = 0
    for step in range(24):
        w.vv(1, dt, s.weight_by_type); sr.dev_vv(1, dt)
        upd_o, tol_o = w.vnl_check()
        tol_g = sr.dev_vnl_check()
        assert abs(tol_g - tol_o) <= 1e-12 * max(tol_o, 1e-30)
        assert sr.vnl_update(tol_g) == upd_o
        if upd_o:
            rebuilds += 1
            w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
            sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
        else:
            assert w.refresh_halo() == 0
            sr.dev_refresh_serial()
        oo, og = w.two_body(), sr.dev_two_body_forces()
        w.vv(2, dt, s.weight_by_type); sr.dev_vv(2, dt)
        assert sr.dev_counts()[1] == w.counts(0)["nlast"]
        assert abs(og[0] - oo[0]) <= 1e-9 * abs(oo[0])
    assert rebuilds >= 2
    sr.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_water_vdw_cutoff_below_rcut_with_exclusions(mode):
    """vdws%cutoff < neigh%cutoff (control.F90:1499-1543) together with exclusion rows: separate vdW / Ewald grids, O-O pairs
    between 6.5 and 8 A carry the Ewald term only, excluded pairs the ewald_excl_forces correction."""
    check_dropin(systems.spce_water(512, rcut=8.0, padding=0.2, rvdw=6.5), 1, mode=mode)
    check_dropin(systems.spce_water(4096, rcut=8.0, padding=0.2, rvdw=6.5), 8, ranks=[2], mode=mode)


def test_frozen_pairs_partition():
    s = systems.nacl(4, rcut=8.0, padding=0.2)
    s.freeze_site[:] = [1, 0]          # all Na+ frozen: Na-Na pairs go to the frozen tail
    s.megfrz = int((s.freeze_site[s.lsite - 1] > 0).sum())
    check_dropin(s, 1)


@pytest.mark.parametrize("mode", [0, 1])
def test_device_list_is_the_reference_list(mode):
    """mode 0: the device rows are the symmetrised reference list; mode 1: the reference's own half rows (as sets)."""
    s = systems.spce_water(512, rcut=8.0, padding=0.2)
    w = world_for(s, P=1)
    d = domain_inputs(w, 0)
    sr = make_sr(s)
    sr.set_force_mode(mode)
    lst = sr.link_cell_pairs(d["natms"], d["nlast"], d["parts"], d["ltype"], d["ltg"], d["lfrzn"], lbook=True, megfrz=0,
                             list_excl=d["list_excl"], max_list=d["max_list"])
    nat = d["natms"]
    main = [set() for _ in range(nat)]
    excl = [set() for _ in range(nat)]
    for i in range(nat):
        n0, n1 = lst[i, 3], lst[i, 2]
        for k in range(1, n1 + 1):
            j = lst[i, 3 + k]
            tgt = main if k <= n0 else excl
            tgt[i].add(j)
            if j <= nat and mode == 0:
                tgt[j - 1].add(i + 1)
    for i in list(range(0, nat, 37)) + [nat - 1]:
        m, x = sr.dev_get_full_row(i + 1)
        assert set(m.tolist()) == main[i] and len(m) == len(main[i])
        assert set(x.tolist()) == excl[i] and len(x) == len(excl[i])
    sr.close()


def test_list_overflow_is_error_106():
    s = systems.argon(6)
    w = world_for(s, P=1)
    d = domain_inputs(w, 0)
    sr = make_sr(s)
    with pytest.raises(engine.DlpError) as ei:
        sr.link_cell_pairs(d["natms"], d["nlast"], d["parts"], d["ltype"], d["ltg"], d["lfrzn"], max_list=20)
    assert ei.value.code == 106 and sr.ibig > 20
    sr.close()


def test_cutoff_too_large_is_error_95():
    s = systems.argon(3)                 # L = 17.2 A < 2 * 8.8
    sr = make_sr(s)
    parts = np.zeros(s.megatm, dtype=COREPART)
    parts["xxx"], parts["yyy"], parts["zzz"] = s.xyz.T
    with pytest.raises(engine.DlpError) as ei:
        sr.link_cell_pairs(s.megatm, s.megatm, parts, np.ones(s.megatm, np.int32), np.arange(1, s.megatm + 1, dtype=np.int32),
                           max_list=s.max_list)
    assert ei.value.code == 95
    sr.close()


def test_two_body_without_list_is_an_error():
    s = systems.argon(6)
    sr = make_sr(s)
    parts = np.zeros(10, dtype=COREPART)
    with pytest.raises(engine.DlpError):
        sr.two_body_forces(10, 10, parts)
    sr.close()


def test_empty_domain():
    s = systems.argon(6)
    sr = make_sr(s)
    parts = np.zeros(0, dtype=COREPART)
    lst = sr.link_cell_pairs(0, 0, parts, np.zeros(0, np.int32), np.zeros(0, np.int32), max_list=s.max_list)
    assert lst.shape[0] == 0
    out = sr.two_body_forces(0, 0, parts)
    assert np.all(out == 0.0)
    sr.close()


# ---- native device-resident path ------------------------------------------------------------------------------------
def native_serial(s):
    sr = engine.ShortRange(0)
    sr.dev_setup_system(s)
    xyz = dd.read_config_fold(s.xyz, s.cell)[0]          # what read_config leaves in parts (the oracle's load does the same)
    sr.dev_load_atoms(xyz, s.vel, np.arange(1, s.megatm + 1, dtype=np.int32), s.lsite)
    return sr


@pytest.mark.parametrize("name", ["argon", "nacl", "water"])
def test_native_serial_halo_list_forces(name):
    s = {"argon": lambda: systems.argon(6), "nacl": lambda: systems.nacl(4, rcut=8.0, padding=0.2),
         "water": lambda: systems.spce_water(512, rcut=8.0, padding=0.2)}[name]()
    w = world_for(s, P=1, with_halo=False, with_list=False)
    w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
    oo = w.two_body()
    sr = native_serial(s)
    sr.dev_relocate_serial()
    sr.dev_halo_serial()
    natms, nlast = sr.dev_counts()
    c = w.counts(0)
    assert (natms, nlast) == (c["natms"], c["nlast"])
    po, pg = w.parts(0), sr.dev_get_parts()
    for k in ("xxx", "yyy", "zzz", "chge"):
        assert np.array_equal(po[k], pg[k]), k                           # halo images: same atoms, same order, same bits
    io, ig = w.ints(0), sr.dev_get_ints()
    for k in ("ltg", "lsite", "ltype", "lfrzn", "ixyz"):
        assert np.array_equal(io[k], ig[k]), k
    sr.dev_link_cell_pairs(want_ref_list=True)
    assert np.array_equal(sr.dev_get_list()[:, :4], w.list(0)[:, :4])
    ref = w.list(0)
    used = np.arange(ref.shape[1] - 4)[None, :] < ref[:, 1:2]
    assert np.array_equal(np.where(used, sr.dev_get_list()[:, 4:], 0), np.where(used, ref[:, 4:], 0))
    out = sr.dev_two_body_forces()
    fg = parts_forces(sr.dev_get_parts(), natms)
    fo = parts_forces(w.parts(0), natms)
    a, b = force_errors(fg, fo)
    assert a <= FORCE_TOL
    assert per_atom_force_error(fg, fo)["per_atom_significant"] <= FORCE_TOL
    for k in range(6):
        assert abs(out[k] - oo[k]) <= ENERGY_TOL * max(abs(oo[k]), 1e-6 * abs(oo[:6]).max())
    sr.close()


def test_native_trajectory_vnl_refresh_rebuild():
    """A short NVE trajectory: vnl_check decisions, refreshed halo positions and rebuilt lists track the oracle."""
    s = systems.argon(6, temperature=300.0)
    dt = 0.004
    w = world_for(s, P=1, with_halo=False, with_list=False)
    sr = native_serial(s)
    w.relocate(); w.set_halo(); w.link_cell_pairs(); w.two_body()
    sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs(); sr.dev_two_body_forces()
    rebuilds = 0
    for step in range(30):
        w.vv(1, dt, s.weight_by_type); sr.dev_vv(1, dt)
        upd_o, tol_o = w.vnl_check()
        tol_g = sr.dev_vnl_check()
        assert abs(tol_g - tol_o) <= 1e-12 * max(tol_o, 1e-30)
        upd_g = sr.vnl_update(tol_g)
        assert upd_g == upd_o
        if upd_o:
            rebuilds += 1
            w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
            sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
        else:
            assert w.refresh_halo() == 0
            sr.dev_refresh_serial()
        oo = w.two_body()
        og = sr.dev_two_body_forces()
        w.vv(2, dt, s.weight_by_type); sr.dev_vv(2, dt)
        natms, nlast = sr.dev_counts()
        assert nlast == w.counts(0)["nlast"]
        po, pg = w.parts(0), sr.dev_get_parts()
        # positions follow the same arithmetic; forces differ by summation order, so trajectories drift at ~1e-13
        assert np.abs(po["xxx"] - pg["xxx"]).max() < 1e-9
        assert abs(og[0] - oo[0]) <= 1e-9 * abs(oo[0])
    assert rebuilds >= 2
    sr.close()


# ---- BASELINE sizes: oracle comparison where it finishes in seconds + size-independent properties --------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c3"])
def test_baseline_configs_full_size(name):
    s = systems.by_name(name)
    w = world_for(s, P=1, with_halo=False, with_list=False)
    w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
    oo = w.two_body()
    sr = native_serial(s)
    sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs(want_ref_list=True)
    natms, nlast = sr.dev_counts()
    ref, got = w.list(0), sr.dev_get_list()
    assert np.array_equal(got[:, :4], ref[:, :4])
    used = np.arange(ref.shape[1] - 4)[None, :] < ref[:, 1:2]
    assert np.array_equal(np.where(used, got[:, 4:], 0), np.where(used, ref[:, 4:], 0))
    out = sr.dev_two_body_forces()
    fg = parts_forces(sr.dev_get_parts(), natms)
    fo = parts_forces(w.parts(0), natms)
    a, b = force_errors(fg, fo)
    assert a <= FORCE_TOL
    rep = per_atom_force_error(fg, fo)
    print("PARITY %s: %s" % (s.name, rep))
    assert rep["per_atom_significant"] <= FORCE_TOL, rep
    for k in range(6):
        assert abs(out[k] - oo[k]) <= ENERGY_TOL * max(abs(oo[k]), 1e-6 * abs(oo[:6]).max())
    # properties that hold at any size: Newton's third law over the periodic system, virial == -trace(stress),
    # symmetric stress
    assert np.abs(fg.sum(0)).max() <= 1e-9 * np.abs(fg).sum()
    vir = out[1] + out[3] + out[5]
    assert abs((out[6] + out[10] + out[14]) + vir) <= 1e-10 * abs(vir)
    assert out[7] == out[9] and out[8] == out[12] and out[11] == out[13]
    sr.close()


@pytest.mark.parametrize("name", ["argon", "nacl_ortho"])
def test_one_kernel_refresh_equals_staged_refresh(name):
    """dlpgpu_dev_refresh_pull (origin map + replayed shifts) gives the staged refresh_halo_positions bit for bit, and both
    equal the oracle's."""
    s = systems.argon(6, temperature=300.0) if name == "argon" else systems.nacl((4, 3, 3), rcut=6.0, padding=0.2, temperature=1200.0)
    w = world_for(s, P=1, with_halo=False, with_list=False)
    w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0; w.two_body()
    sr = native_serial(s)
    sr.dev_p2p_init(0, 1, s.megatm + 64)
    sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs(); sr.dev_two_body_forces()
    for _ in range(2):
        w.vv(1, 0.002, s.weight_by_type); sr.dev_vv(1, 0.002)
        assert w.refresh_halo() == 0
        sr.dev_publish()
        sr.dev_refresh_serial()
        a = sr.dev_get_parts()
        # scramble the halo, then pull
        sr.dev_refresh_pull()
        b = sr.dev_get_parts()
        po = w.parts(0)
        for k in ("xxx", "yyy", "zzz", "chge"):
            assert np.array_equal(a[k], b[k]), k
        natms = s.megatm
        assert np.abs(po["xxx"][natms:] - b["xxx"][natms:]).max() < 1e-9
        w.two_body(); sr.dev_two_body_forces()
        w.vv(2, 0.002, s.weight_by_type); sr.dev_vv(2, 0.002)
    sr.close()


@pytest.mark.parametrize("name", ["argon", "nacl_ortho"])
def test_fused_device_exchange_equals_staged_serial(name):
    """dlpgpu_dev_xchg_rebuild (migration + halo build with device-resident counts, one host sync) leaves exactly the atoms,
    order, tags and checkpoint the staged single-domain routines leave, over a trajectory with atoms crossing the periodic
    boundary; dlpgpu_dev_xchg_gmax returns vnl_check's tolerance."""
    s = systems.argon(6, temperature=300.0) if name == "argon" else systems.nacl((4, 3, 3), rcut=6.0, padding=0.2, temperature=1200.0)
    a, b = native_serial(s), native_serial(s)
    for sr in (a, b):
        sr.set_force_mode(0)      # full list, no atomics: forces are bitwise reproducible, so both trajectories stay identical
    b.dev_p2p_init(0, 1, s.megatm + 64)
    cap_r, cap_h = dd.exchange_capacities(s, (1, 1, 1))
    b.dev_xchg_init(0, 1, cap_r, cap_h)
    neigh = [0] * 6
    a.dev_relocate_serial(); a.dev_halo_serial()
    assert b.dev_xchg_rebuild(neigh, 1) == a.dev_counts()
    seq = 1
    for step in range(6):
        for sr in (a, b):
            sr.dev_link_cell_pairs(); sr.dev_two_body_forces()
        pa, pb = a.dev_get_parts(), b.dev_get_parts()
        ia, ib = a.dev_get_ints(), b.dev_get_ints()
        for k in ("xxx", "yyy", "zzz", "chge", "fxx", "fyy", "fzz"):
            assert np.array_equal(pa[k], pb[k]), (step, k)
        for k in ("ltg", "lsite", "ltype", "lfrzn", "ixyz"):
            assert np.array_equal(ia[k], ib[k]), (step, k)
        assert a.dev_halo_stage_counts() == b.dev_halo_stage_counts()
        for sr in (a, b):
            sr.dev_vv(1, 0.004)
        seq += 1
        assert b.dev_xchg_gmax(seq) == a.dev_vnl_check()
        a.dev_relocate_serial(); a.dev_halo_serial()
        assert b.dev_xchg_rebuild(neigh, seq) == a.dev_counts()
        for sr in (a, b):
            sr.dev_vv(2, 0.004)
    a.close(); b.close()


def test_fused_exchange_reports_a_too_small_stage_buffer():
    """A halo stage that does not fit its receive buffer is error 54, as in export_atomic_data (deport_data.F90:1869-1877)."""
    from dl_poly_b200.lib import DlpError
    s = systems.argon(6, temperature=100.0)
    sr = native_serial(s)
    sr.dev_p2p_init(0, 1, s.megatm + 64)
    sr.dev_xchg_init(0, 1, 64, 16)
    with pytest.raises(DlpError) as e:
        sr.dev_xchg_rebuild([0] * 6, 1)
    assert e.value.code == 54
    sr.close()


def _native_against_oracle(s, P_oracle=1, check_list=True, nthreads=1):
    """The native single-GPU path on system ``s`` against the oracle: resident atoms, the whole reference-format list (P_oracle
    == 1), per-atom forces (by global id), the six energy / virial sums and the stress.  Returns the error report."""
    w = world_for(s, P=P_oracle, with_halo=False, with_list=False)
    w.relocate(); w.set_halo(); assert w.link_cell_pairs(nthreads) == 0
    w.two_body(nthreads)
    oo = sum(w.results(r) for r in range(P_oracle))          # gsum over the oracle's domains (two_body.F90:729)
    sr = native_serial(s)
    sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs(want_ref_list=check_list and P_oracle == 1)
    natms, nlast = sr.dev_counts()
    assert natms == s.megatm
    if check_list and P_oracle == 1:
        ref, got = w.list(0), sr.dev_get_list()
        assert np.array_equal(got[:, :4], ref[:, :4])
        used = np.arange(ref.shape[1] - 4)[None, :] < ref[:, 1:2]
        assert np.array_equal(np.where(used, got[:, 4:], 0), np.where(used, ref[:, 4:], 0))
        del ref, got, used
    # the pair count of the reference's half list, summed over the oracle's domains: local-local pairs once, local-halo pairs
    # on both sides -- for one domain it is the device list's own count
    if P_oracle == 1:
        assert sr.dev_list_pairs() == int(w.list(0)[:, 1].sum())
    out = sr.dev_two_body_forces()
    pg = sr.dev_get_parts()[:natms]
    gid = sr.dev_get_ints()["ltg"][:natms] - 1
    fg = np.zeros((natms, 3))
    fg[gid] = np.stack([pg["fxx"], pg["fyy"], pg["fzz"]], 1)
    fo = w.gather_forces()
    rep = per_atom_force_error(fg, fo)
    scale = np.abs(oo[:6]).max()
    rep["energy_rel"] = float(max(abs(out[k] - oo[k]) / max(abs(oo[k]), 1e-6 * scale) for k in range(6)))
    rep["stress_rel"] = float(np.abs(out[6:15] - oo[6:15]).max() / np.abs(oo[6:15]).max())
    print("PARITY %s: %s" % (s.name, rep))
    assert rep["per_atom_significant"] <= FORCE_TOL, rep        # north star: per-atom forces within 1e-9 relative
    assert rep["max_normalised"] <= FORCE_TOL, rep
    assert rep["energy_rel"] <= ENERGY_TOL, rep                  # engsrp / engcpe / virial terms within 1e-10 relative
    assert rep["stress_rel"] <= ENERGY_TOL, rep
    # size-independent properties: Newton's third law over the periodic box, virial == -trace(stress), symmetric stress
    assert np.abs(fg.sum(0)).max() <= 1e-9 * np.abs(fg).sum()
    vir = out[1] + out[3] + out[5]
    assert abs((out[6] + out[10] + out[14]) + vir) <= 1e-10 * abs(vir)
    assert out[7] == out[9] and out[8] == out[12] and out[11] == out[13]
    sr.close()
    return rep


def test_c4_table_1m_against_oracle():
    """BASELINE configs[3] at its per-GPU size, the size bench.py times (1,000,000 ions, TABLE-file vdW + real-space Ewald):
    the CUDA path against the ORACLE -- the whole reference-format neighbour list bit for bit (counters, members, order),
    per-atom forces, the energy / virial sums and the stress."""
    _native_against_oracle(systems.by_name("c4"))


def test_bench_workload_ionic_1m_against_oracle():
    """The default bench.py workload (molten NaCl, BHM tabulated + Ewald, 1,000,000 ions, seed 1005) against the oracle."""
    _native_against_oracle(systems.by_name("ionic_1m"))


def test_c5_ionic_8m_against_oracle():
    """BASELINE configs[4]: the 8,000,000-ion box on ONE GPU against the oracle's 2x2x2-domain world (its domains run on host
    threads; forces compared by global id, sums after the oracle's gsum)."""
    import os
    try:
        mem_gb = os.sysconf("SC_PAGE_SIZE") * os.sysconf("SC_PHYS_PAGES") / 2.0 ** 30
    except (ValueError, OSError):
        mem_gb = 0.0
    if mem_gb < 56.0:
        pytest.skip("the oracle's eight 1M-ion domains need ~40 GB of host memory")
    _native_against_oracle(systems.by_name("c5_ionic"), P_oracle=8, check_list=False, nthreads=min(8, os.cpu_count() or 1))


def test_c4_table_1m_fast_and_general_pair_kernels_against_each_other():
    """The two pair kernels at 1 M ions: k_pair_v2 (quadratic table form, fp32-completed second differences through the texture
    path; the automatic choice here) and the general kernel that follows the reference statement by statement.  Each is held
    against the oracle in the tests above; this pins them to each other and checks which one actually ran."""
    s = systems.by_name("c4")
    res = []
    for which, expect in ((0, 2), (1, 1)):
        sr = native_serial(s)
        sr.set_pair_kernel(which=which)
        sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
        out = sr.dev_two_body_forces()
        used, _ = sr.pair_kernel_used()
        assert used == expect, (which, used)
        natms, _ = sr.dev_counts()
        res.append((out, parts_forces(sr.dev_get_parts(), natms)))
        sr.close()
    og, fg = res[1]
    of, ff = res[0]
    rep = per_atom_force_error(ff, fg)
    assert rep["per_atom_significant"] <= FORCE_TOL and rep["max_normalised"] <= FORCE_TOL, rep
    for k in range(4):
        assert abs(of[k] - og[k]) <= ENERGY_TOL * abs(og[k]), (k, of[k], og[k])
    for k in range(6, 15):
        assert abs(of[k] - og[k]) <= ENERGY_TOL * np.abs(og[6:15]).max(), (k, of[k], og[k])


@pytest.mark.parametrize("name", ["argon", "nacl", "water", "nacl_ewald_only"])
def test_fast_pair_kernel_is_used_and_matches(name):
    """k_pair_v2 is what runs for the tabulated BASELINE force fields (argon: vdW only; NaCl: vdW + Ewald; SPC/E: with exclusion
    rows; Ewald only), and forcing the general kernel gives the same forces / sums within the bars."""
    s = {"argon": lambda: systems.argon(6), "nacl": lambda: systems.nacl(4, rcut=8.0, padding=0.2),
         "water": lambda: systems.spce_water(512, rcut=8.0, padding=0.2),
         "nacl_ewald_only": lambda: systems.nacl(4, rcut=8.0, padding=0.2, vdw_pairs=())}[name]()
    outs = []
    for which in (0, 1):
        sr = native_serial(s)
        sr.set_pair_kernel(which=which)
        sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
        out = sr.dev_two_body_forces()
        outs.append((sr.pair_kernel_used()[0], out, parts_forces(sr.dev_get_parts(), sr.dev_counts()[0])))
        sr.close()
    assert [o[0] for o in outs] == [2, 1]
    used, out, f = outs[0]
    assert per_atom_force_error(f, outs[1][2])["per_atom_significant"] <= FORCE_TOL
    for k in range(6):
        assert abs(out[k] - outs[1][1][k]) <= ENERGY_TOL * max(abs(outs[1][1][k]), 1e-6 * np.abs(outs[1][1][:6]).max())


@pytest.mark.parametrize("name", ["argon", "nacl", "ionic_mixture", "argon_triclinic", "water"])
def test_list_kernels_build_the_same_rows(name):
    """The half-list kernel with its x-runs trimmed against the cells' bounding boxes (the default), untrimmed (1), and trimmed
    with the per-candidate prune and the shared-memory ring on top (2): identical device rows (length, members AND order) for
    every sampled local atom, and the same pair count."""
    s = {"argon": lambda: systems.argon(8), "nacl": lambda: systems.nacl(6, rcut=8.0, padding=0.2),
         "ionic_mixture": lambda: systems.ionic_mixture(5), "argon_triclinic": lambda: systems.argon_triclinic(6),
         "water": lambda: systems.spce_water(512, rcut=8.0, padding=0.2)}[name]()
    rows = []
    for which in (0, 1, 2):
        sr = native_serial(s)
        sr.set_list_kernel(which)
        sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
        natms, _ = sr.dev_counts()
        rows.append((sr.dev_list_pairs(), [sr.dev_get_full_row(i) for i in range(1, natms + 1, max(1, natms // 300))]))
        sr.close()
    assert rows[0][0] == rows[1][0] == rows[2][0] and rows[0][0] > 0
    for other in (rows[1], rows[2]):
        for ra, rb in zip(rows[0][1], other[1]):
            for xa, xb in zip(ra, rb):
                assert np.array_equal(np.asarray(xa), np.asarray(xb))


@pytest.mark.parametrize("which,P", [("nacl", 1), ("nacl", 8), ("water", 1), ("nacl_frozen", 1), ("nacl_frozen", 8)])
def test_rdf_collect_counts_are_exact(which, P):
    """rdf_collect / rdf_excl_collect / rdf_frzn_collect (rdfs.F90:146-212, :880-946, :948-1018) on the device list: integer pair
    counts per (bin, type pair), summed over the domains, equal the oracle's exactly (bin index from the reference's IEEE
    distance).  nacl_frozen: every Na+ frozen, so the Na-Na pairs sit in the frozen-frozen rows and reach the histogram only
    through rdf_frzn_collect."""
    s = systems.nacl(8 if P > 1 else 4, rcut=8.0, padding=0.2) if which.startswith("nacl") else systems.spce_water(512, rcut=8.0, padding=0.2)
    if which == "nacl_frozen":
        s.freeze_site[:] = [1, 0]
        s.megfrz = int((s.freeze_site[s.lsite - 1] > 0).sum())
    w = world_for(s, P=P)
    nt = s.ff.ntypes
    nkey = nt * (nt + 1) // 2
    rdf_list = np.arange(1, nkey + 1, dtype=np.int32)       # every type pair collected, kk = key
    if which == "water":
        rdf_list[1] = 0                                     # ... except O-H
    n_pairs, max_grid = nkey, 250
    ref = w.rdf_collect(rdf_list, n_pairs, max_grid)
    got = np.zeros((n_pairs, max_grid))
    for r in range(P):
        d = domain_inputs(w, r)
        sr = make_sr(s, d["dd"])
        sr.set_force_mode(1)
        parts = d["parts"].copy()
        sr.link_cell_pairs(d["natms"], d["nlast"], parts, d["ltype"], d["ltg"], d["lfrzn"], lbook=s.lbook, megfrz=s.megfrz,
                           list_excl=d["list_excl"], max_list=d["max_list"], want_list=False)
        sr.two_body_forces(d["natms"], d["nlast"], parts)
        sr.rdf_collect(nt, rdf_list, n_pairs, max_grid, got)
        sr.close()
    assert ref.sum() > 1000
    assert np.array_equal(got, ref)
    if which == "nacl_frozen":
        assert ref[0].sum() > 100          # key 1 = Na-Na: collected although no force row holds such a pair


def test_md_step_enqueued_from_c_follows_the_python_driver():
    """dlpgpu_dev_md_step (the whole step enqueued by the library, sums collected one step late) against Domain.step driven
    from Python: same rebuild decisions, same atom counts, energies equal up to the order of the atomic adds."""
    s = systems.nacl((4, 3, 3), rcut=6.0, padding=0.2, temperature=1200.0)
    a, b = dd.Domain(s, device=0), dd.Domain(s, device=0)
    for d in (a, b):
        d.rebuild(); d.forces()
    eager, lazy, reb_a, reb_b = [], [], [], []
    for step in range(14):
        r0 = a.rebuilds; eager.append(a.step(0.002)); reb_a.append(a.rebuilds != r0)
        r0 = b.rebuilds; prev = b.step(0.002, lazy=True); reb_b.append(b.rebuilds != r0)
        if step > 0:
            lazy.append(prev)
        assert a.sr.dev_counts() == b.sr.dev_counts(), step
    lazy.append(b.collect())
    assert reb_a == reb_b and any(reb_a)
    for k, (x, y) in enumerate(zip(eager, lazy)):
        assert np.allclose(x[:6], y[:6], rtol=1e-9, atol=1e-9 * np.abs(x[:6]).max()), (k, x[:6], y[:6])
    a.close(); b.close()


@pytest.mark.parametrize("key", range(1, 25))
def test_vdw_direct_all_analytic_forms_against_oracle(key):
    """vdw_method direct with each of the 24 analytic forms of two_body_potentials.F90 (the keys of the reference's own
    known-answer list, unit_tests/test_vdw.F90:46-59, which pins the oracle's forms in tests/test_oracle_kat.py): per-atom
    forces, energy / virial and stress of an 864-atom fluid against the oracle."""
    s = systems.vdw_direct_fluid(key)
    rep = _native_against_oracle(s, check_list=False)
    assert rep["significant_atoms"] > 0


def test_vdw_direct_refuses_keys_without_an_analytic_form():
    """A TABLE-file potential (key 0) or an unknown key has nothing vdw_forces_direct could evaluate: dlpgpu_set_vdw says so
    instead of returning zero forces."""
    from dl_poly_b200.lib import DlpError
    s = systems.vdw_direct_fluid(2)
    for bad in (0, 25):
        s.ff.ltp[0] = bad
        sr = engine.ShortRange(0)
        with pytest.raises(DlpError):
            sr.dev_setup_system(s)
        sr.close()


@pytest.mark.parametrize("name", ["argon_direct", "nacl", "water", "morse_direct"])
def test_collect_pp_per_particle_energy_and_stress(name):
    """stats%collect_pp (f3): per-particle energy and stress of the pair terms against the oracle, which restates the reference path
    by path (vdw_forces_direct books the pair energy for every pair, vdw_forces_tab only where the rank owns it, ewald_real for
    every pair; vdw.F90:1707, :1741-1755, :1905, :1987-2001, ewald_spole.F90:155, :205-215).  Also the sum rule: the per-particle
    stresses add up to the stress tensor of the call plus the unowned halo halves."""
    s = {"argon_direct": lambda: systems.argon(6, direct=True), "nacl": lambda: systems.nacl(4, rcut=8.0, padding=0.2),
         "water": lambda: systems.spce_water(512, rcut=8.0, padding=0.2),
         "morse_direct": lambda: systems.vdw_direct_fluid(8)}[name]()
    w = world_for(s, P=1, with_halo=False, with_list=False)
    w.set_collect_pp(True)
    w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
    oo = w.two_body()
    eo, so = w.pp(0)
    sr = native_serial(s)
    sr.set_collect_pp(True)
    sr.dev_relocate_serial(); sr.dev_halo_serial(); sr.dev_link_cell_pairs()
    out = sr.dev_two_body_forces()
    assert sr.pair_kernel_used()[0] == 1
    natms, _ = sr.dev_counts()
    eg, sg = sr.get_pp(natms)
    for k in range(6):
        assert abs(out[k] - oo[k]) <= ENERGY_TOL * max(abs(oo[k]), 1e-6 * np.abs(oo[:6]).max())
    assert np.abs(eg - eo).max() <= 1e-10 * np.abs(eo).max()
    assert np.abs(sg - so).max() <= 1e-10 * np.abs(so).max()
    # the direct and Ewald paths give every local atom half of each of its pairs: the halves add up to the totals
    if name.endswith("direct"):
        assert abs(eg.sum() - (oo[0] + oo[2])) <= 1e-10 * abs(oo[0] + oo[2])
    # ... and get_pp ADDS into the caller's arrays
    e2, s2 = sr.get_pp(natms, eg.copy(), sg.copy())
    assert np.allclose(e2, 2.0 * eg, rtol=1e-15, atol=0.0) and np.allclose(s2, 2.0 * sg, rtol=1e-15, atol=0.0)
    sr.close()


@pytest.mark.parametrize("name,nspl", [("nacl", 8), ("nacl", 6), ("water", 8), ("ionic_mixture", 10), ("nacl_orthorhombic", 8)])
def test_spme_reciprocal_space_against_oracle(name, nspl):
    """f4, first version: ewald_spme_forces_coul on one domain (B-spline spreading, cuFFT, the reference's influence function
    inside its spherical k cutoff, gather with the net force removed, self interaction) against oracle/spme_oracle.py -- the
    numpy restatement that tests/test_spme_oracle.py pins to the exact Ewald reciprocal sum.  Grid and alpha as control.F90:1707-1713
    derives them from spme_precision 1e-6."""
    from oracle import spme_oracle as so
    s = {"nacl": lambda: systems.nacl(4, rcut=8.0, padding=0.2), "water": lambda: systems.spce_water(512, rcut=8.0, padding=0.2),
         "ionic_mixture": lambda: systems.ionic_mixture(5), "nacl_orthorhombic": lambda: systems.nacl((5, 4, 3), rcut=8.0, padding=0.2)}[name]()
    xyz = dd.read_config_fold(s.xyz, s.cell)[0]
    q = s.charge_site[s.lsite - 1]
    _, kdim = so.spme_grid(1.0e-6, s.rcut, s.cell)
    ref = so.ewald_spme_forces_coul(s.cell, xyz, q, s.ff.alpha, kdim, nspl, s.ff.scaling)
    sr = native_serial(s)
    sr.set_spme(kdim, nspl)
    out = sr.dev_spme_forces(s.megatm)
    f = parts_forces(sr.dev_get_parts(), s.megatm)
    assert abs(out[11] - ref["eng_recip"]) <= 1e-10 * abs(ref["eng_recip"]), (out[11], ref["eng_recip"])
    assert abs(out[0] - ref["engcpe_rc"]) <= 1e-10 * abs(ref["engcpe_rc"])
    assert abs(out[1] - ref["vircpe_rc"]) <= 1e-10 * abs(ref["vircpe_rc"])
    assert np.abs(out[2:11] - ref["stress"]).max() <= 1e-10 * np.abs(ref["stress"]).max()
    rep = per_atom_force_error(f, ref["forces"])
    print("SPME %s order %d grid %s: %s" % (s.name, nspl, kdim, rep))
    assert rep["max_normalised"] <= FORCE_TOL and rep["per_atom_significant"] <= FORCE_TOL, rep
    # a second call ADDS the same forces again (every provider adds, drivers.F90:655-660)
    sr.dev_spme_forces(s.megatm)
    f2 = parts_forces(sr.dev_get_parts(), s.megatm)
    assert np.abs(f2 - 2.0 * f).max() <= 1e-12 * np.abs(f).max()
    sr.close()


def test_spme_triclinic_cell_keeps_the_complex_transforms():
    """Orthogonal cells take real-to-complex transforms (half the spectrum); in a parallelepiped cell the reference's influence function
    is not even in m on the Nyquist planes (it takes + K / 2 for both members of a conjugate pair, ewald_spole.F90:1305-1320), so the
    library keeps the complex transforms there.  A sheared NaCl cell against the numpy oracle."""
    from oracle import spme_oracle as so
    base = systems.nacl(4, rcut=8.0, padding=0.2)
    M = np.array([[1.0, 0.20, 0.10], [0.0, 1.0, 0.15], [0.0, 0.0, 1.0]])
    s = systems.System("nacl-sheared", (base.cell.reshape(3, 3) @ M).reshape(9), base.xyz @ M, base.lsite, base.type_site, base.charge_site,
                       base.weight_site, base.ff, base.rcut, base.padding)
    assert s.imcon == 3
    xyz = dd.read_config_fold(s.xyz, s.cell)[0]
    q = s.charge_site[s.lsite - 1]
    _, kdim = so.spme_grid(1.0e-6, s.rcut, s.cell)
    ref = so.ewald_spme_forces_coul(s.cell, xyz, q, s.ff.alpha, kdim, 8, s.ff.scaling)
    sr = native_serial(s)
    sr.set_spme(kdim, 8)
    out = sr.dev_spme_forces(s.megatm)
    f = parts_forces(sr.dev_get_parts(), s.megatm)
    assert abs(out[0] - ref["engcpe_rc"]) <= 1e-10 * abs(ref["engcpe_rc"]) and abs(out[1] - ref["vircpe_rc"]) <= 1e-10 * abs(ref["vircpe_rc"])
    assert np.abs(out[2:11] - ref["stress"]).max() <= 1e-10 * np.abs(ref["stress"]).max()
    rep = per_atom_force_error(f, ref["forces"])
    assert rep["max_normalised"] <= FORCE_TOL and rep["per_atom_significant"] <= FORCE_TOL, rep
    sr.close()


def test_spme_dropin_adds_into_the_callers_records():
    """dlpgpu_spme_forces (host corePart array): the reciprocal forces are ADDED to what parts%f holds, positions and charges
    come back untouched, the sums equal the device-resident call's."""
    from oracle import spme_oracle as so
    s = systems.nacl(4, rcut=8.0, padding=0.2)
    xyz = dd.read_config_fold(s.xyz, s.cell)[0]
    q = s.charge_site[s.lsite - 1]
    _, kdim = so.spme_grid(1.0e-6, s.rcut, s.cell)
    ref = so.ewald_spme_forces_coul(s.cell, xyz, q, s.ff.alpha, kdim, 8, s.ff.scaling)
    parts = np.zeros(s.megatm, dtype=COREPART)
    parts["xxx"], parts["yyy"], parts["zzz"], parts["chge"] = xyz[:, 0], xyz[:, 1], xyz[:, 2], q
    rng = np.random.default_rng(2)
    f0 = rng.standard_normal((s.megatm, 3)) * 1.0e3
    parts["fxx"], parts["fyy"], parts["fzz"] = f0[:, 0], f0[:, 1], f0[:, 2]
    before = parts.copy()
    sr = make_sr(s)
    sr.set_spme(kdim, 8)
    out = sr.spme_forces(s.megatm, parts, s.megatm)
    f = parts_forces(parts, s.megatm) - f0
    assert per_atom_force_error(f, ref["forces"])["max_normalised"] <= FORCE_TOL
    assert abs(out[0] - ref["engcpe_rc"]) <= 1e-10 * abs(ref["engcpe_rc"]) and abs(out[1] - ref["vircpe_rc"]) <= 1e-10 * abs(ref["vircpe_rc"])
    for k in ("xxx", "yyy", "zzz", "chge"):
        assert np.array_equal(parts[k], before[k])
    sr.close()


@pytest.mark.parametrize("threads", [0, 1, 3])
def test_dropin_record_transfer_modes(threads):
    """csrc/hostio.cu: whole records by DMA (0 host threads, the default) or packed fields -- x, y, z, chge up, this provider's
    forces back, added into parts%f by the library's host threads.  Both give the same forces and sums; positions / charges of
    the caller's records are never written; halo records keep their forces.  Packed mode only: whatever another provider
    (tersoff, three-body ...: drivers.F90:675-700) put into parts%f between link_cell_pairs and two_body_forces stays.  The int
    arrays of a second link_cell_pairs call are re-sent only where they changed."""
    s = systems.nacl((6, 6, 6), rcut=8.0, padding=0.3, temperature=1200.0)
    w = world_for(s, P=1)
    w.two_body()
    d = domain_inputs(w, 0)
    kw = dict(lbook=s.lbook, megfrz=s.megfrz, list_excl=d["list_excl"], max_list=d["max_list"])
    natms, nlast = d["natms"], d["nlast"]
    up_rec, down_rec = (64, 64) if threads == 0 else (24, 24)      # packed: x, y, z up (+ 8 for the charge where it changed)
    up_first = 64 if threads == 0 else 32
    rng = np.random.default_rng(7)
    sr = make_sr(s, d["dd"])
    sr.set_host_threads(threads)
    zero = d["parts"].copy()
    for k in ("fxx", "fyy", "fzz"):
        zero[k] = 0.0
    sr.transfer_bytes(reset=True)
    sr.link_cell_pairs(natms, nlast, zero, d["ltype"], d["ltg"], d["lfrzn"], want_list=False, **kw)
    n_int = 2 + (d["lfrzn"] is not None)
    assert sr.transfer_bytes()[0] == up_first * nlast + 4 * n_int * nlast
    sr.transfer_bytes(reset=True)
    out0 = sr.two_body_forces(natms, nlast, zero, unchanged_since_list=True)
    assert sr.transfer_bytes() == (0, down_rec * natms)             # coordinates reused, only the results came back
    f0 = parts_forces(zero, natms)
    fo = parts_forces(d["parts"], natms)
    assert force_errors(f0, fo)[0] <= FORCE_TOL
    other = d["parts"].copy()
    for k in ("fxx", "fyy", "fzz"):
        other[k] = rng.normal(size=nlast) * 1.0e3
    before = other.copy()
    sr.transfer_bytes(reset=True)
    sr.link_cell_pairs(natms, nlast, other, d["ltype"], d["ltg"], d["lfrzn"], want_list=False, **kw)
    assert sr.transfer_bytes()[0] == up_rec * nlast                  # same int arrays as before: none of them sent again
    if threads > 0:
        for k in ("fxx", "fyy", "fzz"):                              # "another provider" adds after the list build
            other[k][:natms] += 17.0
            before[k][:natms] += 17.0
    out1 = sr.two_body_forces(natms, nlast, other, unchanged_since_list=True)
    delta = np.stack([other[k][:natms] - before[k][:natms] for k in ("fxx", "fyy", "fzz")], 1)
    assert np.abs(delta - f0).max() <= 1e-9 * np.abs(f0).max()       # f_before + f rounds once at |f_before| ~ 1e3
    assert np.allclose(out1, out0, rtol=1e-12, atol=1e-12 * np.abs(out0).max())
    for k in ("fxx", "fyy", "fzz"):
        assert np.array_equal(other[k][natms:], before[k][natms:])   # halo records: forces untouched
    for k in ("xxx", "yyy", "zzz", "chge"):
        assert np.array_equal(other[k], before[k])
    # plain call (own upload)
    sr.transfer_bytes(reset=True)
    again = zero.copy()
    for k in ("fxx", "fyy", "fzz"):
        again[k] = 0.0
    sr.two_body_forces(natms, nlast, again)
    assert sr.transfer_bytes() == (up_rec * nlast, down_rec * natms)
    assert np.abs(parts_forces(again, natms) - f0).max() <= 1e-12 * np.abs(f0).max()
    # vnl_check sends parts(1:natms), the force call parts(1:nlast): neither re-sends a charge
    sr.transfer_bytes(reset=True)
    assert sr.vnl_check(natms, again) == 0.0
    assert sr.transfer_bytes()[0] == up_rec * natms
    sr.transfer_bytes(reset=True)
    sr.two_body_forces(natms, nlast, again)
    assert sr.transfer_bytes()[0] == up_rec * nlast
    for k in ("fxx", "fyy", "fzz"):
        again[k] = 0.0
    # one atom of another type: the chunk that holds it is sent again and the forces follow
    lt = d["ltype"].copy()
    lt[5] = lt[5] % 2 + 1
    sr.transfer_bytes(reset=True)
    sr.link_cell_pairs(natms, nlast, again, lt, d["ltg"], d["lfrzn"], want_list=False, **kw)
    sent = sr.transfer_bytes()[0] - up_rec * nlast
    assert 0 < sent <= 4 * nlast
    for k in ("fxx", "fyy", "fzz"):
        again[k] = 0.0
    sr.two_body_forces(natms, nlast, again, unchanged_since_list=True)
    assert np.abs(parts_forces(again, natms)[5] - f0[5]).max() > 1e-6 * np.abs(f0).max()
    # a changed charge travels with its chunk
    if threads > 0:
        q = again.copy()
        q["chge"][7] *= 0.5
        for k in ("fxx", "fyy", "fzz"):
            q[k] = 0.0
        sr.transfer_bytes(reset=True)
        sr.two_body_forces(natms, nlast, q)
        sent = sr.transfer_bytes()[0] - up_rec * nlast
        assert 0 < sent <= 8 * nlast
        assert np.abs(parts_forces(q, natms)[7]).max() > 0 and not np.allclose(parts_forces(q, natms)[7], parts_forces(again, natms)[7])
    # switching the mode invalidates what the device holds of the caller's records
    sr.set_host_threads(2 if threads == 0 else 0)
    with pytest.raises(Exception):
        sr.two_body_forces(natms, nlast, again, unchanged_since_list=True)
    for k in ("fxx", "fyy", "fzz"):
        zero[k] = 0.0
    sr.link_cell_pairs(natms, nlast, zero, d["ltype"], d["ltg"], d["lfrzn"], want_list=False, **kw)
    sr.two_body_forces(natms, nlast, zero, unchanged_since_list=True)
    assert np.abs(parts_forces(zero, natms) - f0).max() <= 1e-12 * np.abs(f0).max()
    sr.close()
