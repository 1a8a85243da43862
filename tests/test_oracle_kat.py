"""Pins the CPU oracle against the reference's own known-answer vectors and an independent brute force (CPU only)."""
import json
import os

import numpy as np
import pytest

from dl_poly_b200 import systems, tables
from util import world_for

GOLD = os.path.join(os.path.dirname(__file__), "golden")

# source/unit_tests/test_vdw.F90:46-59 (literals), fixture :93-177: r = 1, dr = (1,1,1), params = 1..7, rvdw = 10
EXPECTED_E = [-1.0, 16128.0, 80.0, -2.3934693402873668, 45.598150033144236, -1.0, 4567.8016528926664, 363.25771964635982,
              1.0, 0.25, 555.50928442334305, 15616.0, 367.25771964635982, 2.1495939317213679, 16882.652704957309,
              3.8142266913481757e30, 11761.890149027380, 16128.0, -2.3934693402873668, -1.0, 4.0, -0.8948393168143698,
              -0.4921875, -1765.997274794514]
EXPECTED_V = [0.0, -195072.0, -288.0, 17.696734670143684, -45.196300066288472, 8.0, -16899.173553719396, -2300.0595394172847,
              12.0, -0.5, -3719.0014773362891, -192000.0, -2348.0595394172847, 0.71653131057378916, -50044.066263312896,
              -5.2445617006037446e31, -36135.103419546060, -195072.0, 17.696734670143684, 0.0, -40.0, -0.1988531815143044,
              -0.09375, 15903.665444480088]


@pytest.mark.parametrize("k", range(1, 25))
def test_vdw_direct_known_answers(oracle, k):
    e, v = oracle.kat_vdw_direct(k, [1, 2, 3, 4, 5, 6, 7])
    # asserts.F90:20 default tolerance is 1e-6 absolute; the restatement agrees to ~1 ulp, so hold it to 1e-13 relative
    assert abs(e - EXPECTED_E[k - 1]) <= 1e-13 * max(1.0, abs(EXPECTED_E[k - 1]))
    assert abs(v - EXPECTED_V[k - 1]) <= 1e-13 * max(1.0, abs(EXPECTED_V[k - 1]))


def test_match_binary_search(oracle):
    lst = [2, 3, 7, 11, 40, 41, 97]
    for n in range(1, 100):
        assert oracle.match(n, lst) == (n in lst)
    assert oracle.match(5, []) is False


def test_dcell_invert_cubic_and_triclinic(oracle):
    cell = np.array([10.0, 0, 0, 0, 12.0, 0, 0, 0, 14.0])
    d = oracle.dcell(cell)
    assert np.allclose(d[:3], [10, 12, 14]) and np.allclose(d[6:9], [10, 12, 14]) and abs(d[9] - 1680) < 1e-9
    tri = np.array([10.0, 0, 0, 3.0, 9.0, 0, 1.0, 2.0, 8.0])
    inv, det = oracle.invert(tri)
    assert np.allclose(inv.reshape(3, 3) @ tri.reshape(3, 3), np.eye(3), atol=1e-14)
    assert abs(det - np.linalg.det(tri.reshape(3, 3))) < 1e-9


def test_images_minimum_image(oracle):
    rng = np.random.default_rng(5)
    x, y, z = (rng.uniform(-40, 40, 50) for _ in range(3))
    cell = np.diag([20.0, 20.0, 20.0]).reshape(9)
    a, b, c = oracle.images(1, cell, x, y, z)
    assert np.all(np.abs(a) <= 10 + 1e-12) and np.all(np.abs(b) <= 10 + 1e-12)
    assert np.allclose((a - x) / 20.0, np.rint((a - x) / 20.0), atol=1e-12)


def test_erfc_tables_against_scipy(oracle):
    from scipy.special import erfc
    rcut, alpha = 12.0, tables.ewald_alpha(1e-6, 12.0)
    assert abs(alpha - 0.26506) < 5e-5
    n = oracle.lib().ora_max_grid(rcut)
    assert n == 1204
    et, dt, rs = oracle.erfcgen(rcut, alpha, n)
    r = np.arange(1, n + 1) * (rcut / (n - 4))
    assert np.max(np.abs(et[1:] * r - erfc(alpha * r))) < 2e-7          # A&S 7.1.26: |error| <= 1.5e-7
    et2, dt2, rs2 = tables.erfcgen(rcut, alpha)
    assert rs == rs2
    assert np.max(np.abs(et - et2) / np.maximum(np.abs(et), 1e-300)) < 1e-14
    assert np.max(np.abs(dt - dt2) / np.maximum(np.abs(dt), 1e-300)) < 1e-14


@pytest.mark.parametrize("form,key,param", [("12-6", 1, [4 * 99.61 * 3.405 ** 12, 4 * 99.61 * 3.405 ** 6]), ("lj", 2, [99.61, 3.405]),
                                           ("buck", 4, [1.0e5, 0.3, 500.0]), ("bhm", 5, [2544.35, 3.1545, 2.34, 1.0117e4, 4.8177e3])])
def test_host_tables_match_oracle(oracle, form, key, param):
    g = tables.max_grid(8.5)
    tp, tf = tables.vdw_generate(key, np.array(param + [0] * (7 - len(param)), dtype=float), 8.5, g)
    op, of = oracle.vdw_generate(key, param, 8.5, g)
    assert tp[0] == op[0] and tf[0] == of[0]                            # Huge()
    # numpy's exp and glibc's differ by <= 1 ulp; the exp-6 forms amplify that through cancellation -> 1e-12
    assert np.max(np.abs(tp[1:] - op[1:]) / np.abs(op[1:])) < 1e-12
    assert np.max(np.abs(tf[1:] - of[1:]) / np.abs(of[1:])) < 1e-12
    a, b = tables.vdw_direct_fs(key, np.array(param + [0] * (7 - len(param)), dtype=float), 8.5)
    oa, ob = oracle.vdw_direct_fs(key, param, 8.5)
    assert abs(a - oa) <= 1e-12 * abs(oa) and abs(b - ob) <= 1e-12 * abs(ob)


def test_table_regrid_identity_and_remake(oracle):
    g = tables.max_grid(8.0)
    dl = 8.0 / (g - 4)
    r = np.arange(1, g + 1) * dl
    e, gm = tables.pot_energy(2, [65.0, 3.166], r)
    same = tables.regrid_table(e, dl, 8.0, g, False)
    assert np.array_equal(same[1:g - 3], e[:g - 4])
    assert np.array_equal(same, oracle.vdw_table_regrid(e, dl, 8.0, g, False))
    # coarser file grid -> 3-point re-gridding path (vdw.F90:1212-1238)
    dl2 = 0.0095
    r2 = np.arange(1, 900) * dl2
    e2, g2 = tables.pot_energy(2, [65.0, 3.166], r2)
    a = tables.regrid_table(g2, dl2, 8.0, g, True)
    b = oracle.vdw_table_regrid(g2, dl2, 8.0, g, True)
    assert np.array_equal(a, b)
    ref = tables.pot_energy(2, [65.0, 3.166], r[300:700])[1]
    assert np.max(np.abs(a[301:701] - ref) / np.abs(ref)) < 1e-3          # 3-point interpolation error, sanity only


def test_table_file_roundtrip(tmp_path):
    g = tables.max_grid(8.0)
    dl = 8.0 / (g - 4)
    r = np.arange(1, g + 1) * dl
    e, gm = tables.pot_energy(5, [2544.35, 3.1545, 2.34, 1.0117e4, 4.8177e3], r)
    p = tmp_path / "TABLE"
    tables.write_table_file(str(p), [("Na+", "Na+", 1.5, -2.5, e, gm)], dl, 8.0, g)
    delpot, cutpot, ngrid, pairs = tables.read_table_file(str(p))
    assert ngrid == g and abs(delpot - dl) < 1e-15 and pairs[0][0] == "Na+"
    assert np.max(np.abs(pairs[0][4] - e) / np.abs(e)) < 1e-15


CASES = {
    "argon_864": lambda: systems.argon(6),
    "nacl_512": lambda: systems.nacl(4, rcut=8.0, padding=0.2),
    "water_1536": lambda: systems.spce_water(512, rcut=8.0, padding=0.2),
    "argon_triclinic_864": lambda: systems.argon_triclinic(6),                       # imcon = 3, parallelepiped cell
    "water_rvdw_below_rcut": lambda: systems.spce_water(512, rcut=8.0, padding=0.2, rvdw=6.5),
}


def _with_pdplnc(s, v):
    s.pdplnc = v
    return s


@pytest.mark.parametrize("pdplnc,nlp", [(2.0, 3), (0.5, 4)])
def test_oracle_nlp3_shortcut_over_admits_but_forces_agree(oracle, pdplnc, nlp):
    """neighbours.F90:537: for nlp >= 3 the cells within (nlp-1)^2 of the primary cell skip the distance test, so the list
    is a SUPERSET of the pairs within cutoff_extended; the force loops test r < rcut themselves, so forces
    and energies still equal the brute force."""
    s = _with_pdplnc(systems.argon(6), pdplnc)
    w = world_for(s, P=1)
    c = w.counts(0)
    assert c["nlp"] == nlp
    out = w.two_body()
    x = w.gather_positions()
    bp, band = oracle.brute_pairs(x, s.cell, s.rx)
    got = set()
    lst, ints = w.list(0), w.ints(0)
    for i in range(c["natms"]):
        for k in range(1, lst[i, 1] + 1):
            j = lst[i, 3 + k] - 1
            gi, gj = ints["ltg"][i], ints["ltg"][j]
            if j < c["natms"] or gi < gj:
                got.add((min(gi, gj), max(gi, gj)))
    want = set(map(tuple, bp))
    assert want <= got
    if nlp == 3:
        assert len(got) > len(want)                          # over-admission is real at this cell size
    fb, ob = w.brute_forces(x, s.lsite)
    f = w.gather_forces()
    assert np.abs(f - fb).max() <= 1e-12 * np.abs(fb).max()
    for k in range(2):
        assert abs(out[k] - ob[k]) <= 1e-12 * max(abs(ob[k]), 1.0)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_vs_brute_force(oracle, name):
    """Second opinion for the parts the reference leaves unpinned: pair set by definition, forces/energies in long double."""
    s = CASES[name]()
    w = world_for(s, P=1)
    out = w.two_body()
    x = w.gather_positions()
    bp, band = oracle.brute_pairs(x, s.cell, s.rx)
    got = set()
    c, lst, ints = w.counts(0), w.list(0), w.ints(0)
    for i in range(c["natms"]):
        for k in range(1, lst[i, 1] + 1):          # list(-2,i) = full row length
            j = lst[i, 3 + k] - 1
            gi, gj = ints["ltg"][i], ints["ltg"][j]
            if j < c["natms"] or gi < gj:
                got.add((min(gi, gj), max(gi, gj)))
    assert band == 0
    assert got == set(map(tuple, bp))
    fb, ob = w.brute_forces(x, s.lsite)
    f = w.gather_forces()
    assert np.abs(f - fb).max() <= 1e-12 * np.abs(fb).max()
    for k in range(6):
        assert abs(out[k] - ob[k]) <= 1e-12 * max(abs(ob[k]), 1.0)
    assert abs((out[6] + out[10] + out[14]) + (out[1] + out[3] + out[5])) <= 1e-10 * abs(out[1] + out[3] + out[5])


def test_oracle_multidomain_equals_serial(oracle):
    """8 domains (2x2x2, map_domains) give the same pair set / forces / energies as the serial run."""
    s = systems.nacl(8, rcut=8.0, padding=0.2)
    w1 = world_for(s, P=1)
    w8 = world_for(s, P=8)
    assert tuple(w8.dd(0)[0][:3]) == (2, 2, 2)
    o1, o8 = w1.two_body(), w8.two_body()
    f1, f8 = w1.gather_forces(), w8.gather_forces()
    assert np.abs(f1 - f8).max() <= 1e-12 * np.abs(f1).max()
    for k in range(15):
        assert abs(o1[k] - o8[k]) <= 1e-11 * max(abs(o1[k]), 1.0)


def test_golden_fixture_matches_oracle(oracle):
    """tests/golden/*.json were produced by tests/golden/make_golden.py from this oracle; they freeze its outputs so a
    later edit of the restatement cannot drift silently."""
    for fn in sorted(os.listdir(GOLD)):
        if not fn.endswith(".json"):
            continue
        g = json.load(open(os.path.join(GOLD, fn)))
        if "generator" not in g or "P" not in g:      # host_logic.json / the SPME fixtures (tests/test_spme_oracle.py) are read elsewhere
            continue
        s = getattr(systems, g["generator"])(**g["kwargs"])
        w = world_for(s, P=g["P"])
        out = w.two_body()
        assert [w.counts(r)["nlast"] for r in range(g["P"])] == g["nlast"]
        assert int(sum(w.list(r)[:, 1].sum() for r in range(g["P"]))) == g["list_entries"]
        for k, v in enumerate(g["out"]):
            assert abs(out[k] - v) <= 1e-13 * max(abs(v), 1.0), (fn, k)
        f = w.gather_forces()
        assert abs(float(np.abs(f).sum()) - g["force_l1"]) <= 1e-12 * g["force_l1"]


def test_read_config_fold_matches_oracle_load_bits(oracle):
    """dd.read_config_fold (numpy) reproduces the oracle's read_config restatement bit for bit, incl. domain assignment."""
    from dl_poly_b200 import dd
    for s, P in ((systems.argon(6), 1), (systems.nacl((4, 2, 2), rcut=5.0, padding=0.2), 4), (systems.spce_water(512, rcut=8.0, padding=0.2), 8)):
        w = oracle.World.from_system(s, P=P)
        dims = dd.map_domains(P, dd.cell_widths(s.cell), s.imcon)
        assert tuple(w.dd(0)[0][:3]) == dims
        xyz, owner = dd.read_config_fold(s.xyz, s.cell, dims)
        got = w.gather_positions()
        assert np.array_equal(got, xyz)
        for r in range(P):
            assert np.array_equal(np.sort(w.ints(r)["ltg"][:w.counts(r)["natms"]]), np.nonzero(owner == r)[0] + 1)
            assert list(w.dd(r)[1][:6]) == dd.face_neighbours(r, *dims)


@pytest.mark.parametrize("kind,eps", [("coul", 1.0), ("dddp", 2.5), ("fscp", 1.0), ("rfp", 78.0)])
def test_direct_coulomb_variants_against_numpy_brute_force(oracle, kind, eps):
    """coul_spole.F90 restatement (coul_cp / coul_dddp / coul_fscp / coul_rfp_forces, undamped) against an independent
    O(N^2) minimum-image evaluation of the closed forms in numpy: energy, virial, per-atom forces."""
    s = systems.nacl(2, rcut=5.0, padding=0.2, coulomb=kind, eps=eps, vdw_pairs=())
    w = world_for(s, P=1)
    out = w.two_body()
    n = s.megatm
    xyz = w.parts(0)[:n]
    pos = np.stack([xyz["xxx"], xyz["yyy"], xyz["zzz"]], 1)
    q = s.charge_site[s.lsite - 1]
    L = s.cell[0]
    d = pos[:, None, :] - pos[None, :, :]
    d -= L * np.rint(d / L)
    r = np.sqrt((d ** 2).sum(-1))
    iu = np.triu_indices(n, 1)
    rr, dd = r[iu], d[iu]
    m = rr < s.rcut
    rr, dd, ii, jj = rr[m], dd[m], iu[0][m], iu[1][m]
    qq = q[ii] * q[jj] * tables.R4PIE0 / eps
    rc = s.rcut
    if kind == "coul":
        e, g = qq / rr, qq / rr ** 3
    elif kind == "dddp":
        e, g = qq / rr ** 2, 2.0 * qq / rr ** 4
    elif kind == "fscp":
        e, g = qq * (1.0 / rr + rr / rc ** 2 - 2.0 / rc), qq * (1.0 / rr ** 2 - 1.0 / rc ** 2) / rr
    else:
        b0 = 2.0 * (eps - 1.0) / (2.0 * eps + 1.0)
        e, g = qq * (1.0 / rr + 0.5 * b0 * rr ** 2 / rc ** 3 - (1.0 + 0.5 * b0) / rc), qq * (1.0 / rr ** 3 - b0 / rc ** 3)
    vir = {"coul": -e.sum(), "dddp": -2.0 * e.sum()}.get(kind, -(g * rr ** 2).sum())
    f = np.zeros((n, 3))
    np.add.at(f, ii, g[:, None] * dd)
    np.add.at(f, jj, -g[:, None] * dd)
    assert abs(out[2] - e.sum()) <= 1e-11 * np.abs(e).sum()
    assert abs(out[3] - vir) <= 1e-11 * np.abs(g * rr ** 2).sum()
    fo = np.stack([xyz["fxx"], xyz["fyy"], xyz["fzz"]], 1)
    assert np.abs(fo - f).max() <= 1e-10 * np.abs(f).max()


def test_rdf_collect_against_numpy_histogram(oracle):
    """rdfs.F90:146-212 restatement against an O(N^2) minimum-image histogram: ll = min(1 + Int(r rdelr), max_grid)."""
    s = systems.nacl(3, rcut=6.0, padding=0.2)
    w = world_for(s, P=1)
    nt = s.ff.ntypes
    rdf_list = np.array([1, 2, 3], dtype=np.int32)
    max_grid = 120
    got = w.rdf_collect(rdf_list, 3, max_grid)
    xyz = w.parts(0)[:s.megatm]
    pos = np.stack([xyz["xxx"], xyz["yyy"], xyz["zzz"]], 1)
    t = s.type_site[s.lsite - 1]
    L = s.cell[0]
    d = pos[:, None, :] - pos[None, :, :]
    d -= L * np.rint(d / L)
    r = np.sqrt((d ** 2).sum(-1))
    iu = np.triu_indices(s.megatm, 1)
    rr = r[iu]
    hi, lo = np.maximum(t[iu[0]], t[iu[1]]), np.minimum(t[iu[0]], t[iu[1]])
    key = hi * (hi - 1) // 2 + lo
    m = rr < s.rcut
    ll = np.minimum(1 + (rr[m] * (max_grid / s.rcut)).astype(np.int64), max_grid)
    ref = np.zeros((3, max_grid))
    np.add.at(ref, (key[m] - 1, ll - 1), 1.0)
    assert got.sum() == ref.sum() == m.sum()
    # a pair sitting within an ulp of a bin edge may fall either way between the two distance evaluations
    assert np.abs(got - ref).sum() <= 4


def test_golden_host_logic_matches_oracle(oracle):
    """tests/golden/host_logic.json: the oracle's vdw_lrc, end-of-two_body_forces and vnl_check decision answers, frozen."""
    import test_host_cpp as th
    g = json.load(open(os.path.join(GOLD, "host_logic.json")))
    s = systems.nacl(3, rcut=9.0, padding=0.2)
    num_type = [float((s.type_site[s.lsite - 1] == t).sum()) for t in (1, 2)]
    e, v = th._ora_lrc(oracle, s.ff, num_type, [3.0, 5.0], 1, s.volume)
    assert [e, v] == g["vdw_lrc_nacl_216_rc9_frozen_3_5"]
    a = systems.argon(4)
    assert list(th._ora_lrc(oracle, a.ff, [float(a.megatm)], [0.0], 1, a.volume)) == g["vdw_lrc_argon_256"]
    tot, st = th._ora_epilogue(oracle, [1.0e5, -2.0e5, -3.0e6, 2.5e6, 4.0e3, -5.0e3, -7.0e5, 6.0e5], True, 2.0, 0.3, 1.0, 5000.0, e, v, 8,
                               np.arange(9.0))
    assert list(tot) == g["epilogue_totals"] and list(st) == g["epilogue_stress"]
    rows, kodes = th._ora_vnl_trace(oracle, False, 0, 8.5, 0.1, np.diag([114.39] * 3).reshape(9), [1, 1, 1], g["vnl_tols"])
    assert kodes == g["vnl_kodes"] and np.array_equal(rows, np.array(g["vnl_trace_nostrict_argon"]))
