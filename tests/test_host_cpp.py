"""The C++ host side above the C ABI (dl-poly_b200/host/dlpoly_host.cpp), driven through its check program.

CPU part (no GPU): dcell / invert / map_domains / vdw_generate / vdw_table_read / erfcgen / coul_setup / the decision logic of
vnl_check against the oracle, bit for bit.  GPU part: one domain through link_cell_pairs + two_body_forces (host buffers, the
call sequence of the Fortran call sites) and a native NVE run, against the oracle with the north-star bars.
"""
import os
import struct
import subprocess

import numpy as np
import pytest

from dl_poly_b200 import build as dlp_build
from dl_poly_b200 import systems, tables
from dl_poly_b200.lib import COREPART
from util import domain_inputs, force_errors, parts_forces, world_for, per_atom_force_error

FORCE_TOL = 1.0e-9      # north star: per-atom forces, relative to the largest force
ENERGY_TOL = 1.0e-10    # north star: energies / virials, relative

ELECTRO_KEY = {0: 0, "spme": 1, "dddp": 2, "coul": 3, "fscp": 4, "rfp": 5}      # electrostatic.F90:18-28


# ---------------------------------------------------------------- bundle files (dlpoly_check.cpp)
def write_bundle(path, recs):
    with open(path, "wb") as f:
        for name, val in recs.items():
            if isinstance(val, (bytes, str)):
                data, dt, n = (val.encode() if isinstance(val, str) else val), b"b", None
                n = len(data)
            else:
                a = np.ascontiguousarray(val)
                if a.dtype.kind == "f":
                    a, dt = a.astype(np.float64), b"d"
                    n = a.size
                elif a.dtype.kind in "iub":
                    a, dt = a.astype(np.int32), b"i"
                    n = a.size
                else:                       # structured records (corePart) travel as bytes
                    dt, n = b"b", a.nbytes
                data = a.tobytes()
            head = name.encode().ljust(24, b"\0") + dt + b"\0" * 7
            f.write(head + struct.pack("<q", n) + data)


def read_bundle(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            head = f.read(32)
            if len(head) < 32:
                break
            n = struct.unpack("<q", f.read(8))[0]
            name = head[:24].split(b"\0")[0].decode()
            dt = head[24:25]
            if dt == b"d":
                out[name] = np.frombuffer(f.read(8 * n), dtype=np.float64).copy()
            elif dt == b"i":
                out[name] = np.frombuffer(f.read(4 * n), dtype=np.int32).copy()
            else:
                out[name] = f.read(n)
    return out


def run_check(mode, recs, tmp_path, expect_rc=0):
    exe = dlp_build.build_host()
    fin, fout = str(tmp_path / ("%s_in.bundle" % mode)), str(tmp_path / ("%s_out.bundle" % mode))
    write_bundle(fin, recs)
    r = subprocess.run([exe, mode, fin, fout], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == expect_rc, (r.returncode, r.stderr)
    return read_bundle(fout) if expect_rc == 0 else r.stderr


def ff_records(s, ff_spec):
    """The FIELD / CONTROL facts the C++ host builds its force field from (ff_spec: [(ai, aj, form, params)])."""
    ff = s.ff
    pairs, prm = [], []
    for ai, aj, form, p in ff_spec:
        pairs += [ai, aj, tables.KEYPOT[form]]
        q = np.zeros(7)
        q[:len(p)] = p
        prm.append(q)
    key = 1 if ff.ew_active else {0: 0, 1: 3, 2: 2, 3: 4, 4: 5}[ff.coul_kind]
    recs = dict(cell=s.cell, imcon=[s.imcon], megatm=[s.megatm], megfrz=[s.megfrz], ntypes=[ff.ntypes], rvdw=[ff.rvdw],
                rcut=[s.rcut], padding=[s.padding], pdplnc=[s.pdplnc], max_list=[s.max_list], force_shift=[int(ff.force_shift)],
                direct=[int(ff.direct)], pot_pairs=np.array(pairs, dtype=np.int32), pot_param=np.array(prm).reshape(-1),
                electro_key=[key], eps=[ff.eps], damping=[ff.alpha if ff.coul_kind else 0.0], fdens=[s.density])
    if not len(pairs):
        recs["pot_pairs"] = np.zeros(0, dtype=np.int32)
        recs["pot_param"] = np.zeros(0)
    if ff.ew_active:
        recs["ew_alpha"] = [ff.alpha]
    return recs


BHM = systems._BHM
SPEC_NACL = [(ai, aj, "bhm", p) for (ai, aj), p in BHM.items()]
EPS_AR, SIG_AR = 99.61, 3.405
SPEC_AR_126 = [(1, 1, "12-6", [4 * EPS_AR * SIG_AR ** 12, 4 * EPS_AR * SIG_AR ** 6])]
SPEC_AR_LJ = [(1, 1, "lj", [EPS_AR, SIG_AR])]
SPEC_WATER = [(1, 1, "lj", [65.0, 3.166])]


def close(a, b, rtol=1.0e-12):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= rtol * np.abs(b)))


def check_ff_equal(out, ff):
    """Every array the C++ host generated against what tables.ForceField holds: integers and sizes exactly, the tables to
    1e-12 (numpy's exp and glibc's differ by up to 1 ulp, which the exp-6 forms amplify); the oracle comparison in the
    caller is the bit-exact one."""
    assert np.array_equal(out["vdw_list"], ff.vdw_list_c)
    assert np.array_equal(out["ltp"], ff.ltp)
    assert tuple(out["vdw_sizes"]) == (ff.n_vdw, ff.max_vdw, ff.mxgrid)
    assert close(out["tab_potential"], ff.tab_potential.reshape(-1))
    assert close(out["tab_force"], ff.tab_force.reshape(-1))
    assert close(out["afs"], ff.afs) and close(out["bfs"], ff.bfs)
    assert np.array_equal(out["param"], ff.param.reshape(-1))
    if ff.ew_active or ff.coul_damp:
        assert close(out["erfc"], ff.erfc) and close(out["erfc_deriv"], ff.erfc_deriv)
        assert out["electro"][1] == ff.ew_recip
    if ff.coul_kind:
        assert close(out["electro"][2:4], [ff.coul_force_shift, ff.coul_energy_shift])
        assert close(out["electro"][4:7], ff.coul_rf, 1e-15)


# ---------------------------------------------------------------- CPU: numerics, domains
def test_cpp_dcell_invert_map_domains_equal_the_oracle(oracle, tmp_path):
    rng = np.random.default_rng(11)
    cells = [np.diag([30.0, 30.0, 30.0]).reshape(9), np.diag([20.0, 35.0, 50.0]).reshape(9),
             np.array([30.0, 0, 0, 4.0, 28.0, 0, 3.0, 5.0, 33.0]), np.array([0.0, 25.0, 0, 31.0, 0, 0, 2.0, 1.0, 40.0])]
    cells += [(np.diag(rng.uniform(20, 60, 3)) + rng.uniform(-4, 4, (3, 3))).reshape(9) for _ in range(4)]
    for cell in cells:
        widths = oracle.dcell(cell)[6:9]
        cases, wid = [], []
        for P in (1, 2, 3, 4, 6, 8, 12, 16, 18, 27, 30, 64):
            for rank in sorted({0, P // 2, P - 1}):
                cases += [1, P, rank]
                wid += list(widths)
        out = run_check("host", dict(cell=cell, dd_cases=np.array(cases, dtype=np.int32), dd_widths=np.array(wid)), tmp_path)
        assert np.array_equal(out["celprp"], oracle.dcell(cell))
        from dl_poly_b200 import dd
        assert np.array_equal(np.array(dd.dcell(cell)), oracle.dcell(cell))          # the Python host's dcell: same bits
        inv, det = oracle.invert(cell)
        assert np.array_equal(out["rcell"], inv) and out["det"][0] == det
        res = out["dd_results"].reshape(-1, 58)
        k = 0
        for P in (1, 2, 3, 4, 6, 8, 12, 16, 18, 27, 30, 64):
            w = oracle.World(P, cell, 1)
            for rank in sorted({0, P // 2, P - 1}):
                dd6, map26 = w.dd(rank)
                assert np.array_equal(res[k, :6], dd6), (P, rank)
                assert np.array_equal(res[k, 6:32], map26), (P, rank)
                uniq = [int(map26[i] == rank or map26[i] in map26[:i]) for i in range(26)]      # domains.F90:251-256
                assert list(res[k, 32:58]) == uniq
                k += 1


def test_cpp_read_config_fold_equals_the_oracle_load(oracle, tmp_path):
    """configuration.F90:1183-1205: folded positions (same bits) and domain assignment as the oracle's read_config restatement."""
    for s, P in ((systems.argon(6), 1), (systems.nacl((4, 2, 2), rcut=5.0, padding=0.2), 4), (systems.spce_water(512, rcut=8.0, padding=0.2), 8)):
        shifted = s.xyz + np.array([3.0, -2.0, 1.0]) * np.array([s.cell[0], s.cell[4], s.cell[8]])   # images far outside the cell
        w = oracle.World(P, s.cell, s.imcon)
        w.set_cutoffs(s.rcut, s.padding, s.pdplnc)
        w.set_sites(s.type_site, s.charge_site, s.freeze_site)
        w.load(shifted, None, s.lsite)
        out = run_check("host", dict(cell=s.cell, imcon=[s.imcon], mxnode=[P], fold_xyz=shifted), tmp_path)
        assert np.array_equal(out["folded_xyz"].reshape(-1, 3), w.gather_positions())
        for r in range(P):
            assert np.array_equal(np.sort(w.ints(r)["ltg"][:w.counts(r)["natms"]]), np.nonzero(out["owner"] == r)[0] + 1)


def test_cpp_exchange_capacities_equal_the_python_engine(tmp_path):
    """Stage-buffer capacities of the device-side exchange must be identical on every rank and in both hosts."""
    from dl_poly_b200 import dd
    cases = [(systems.nacl((8, 8, 8), rcut=8.0, padding=0.3), 8), (systems.nacl((6, 4, 4), rcut=6.0, padding=0.2), 4),
             (systems.argon(8), 2), (systems.argon(6), 1), (systems.spce_water(512, rcut=8.0, padding=0.2), 1)]
    for s, P in cases:
        dims = dd.map_domains(P, dd.cell_widths(s.cell), s.imcon)
        want = dd.exchange_capacities(s, dims, safety=2.0)
        out = run_check("host", dict(cell=s.cell, imcon=[s.imcon], cap_cases=np.array([P, s.megatm], dtype=np.int32),
                                     cap_cutoffs=np.array([s.rcut, s.padding])), tmp_path)
        assert tuple(out["cap_results"]) == tuple(want), (s.name, P)


def test_cpp_map_domains_slab_limits(tmp_path):
    """imcon 0 (no periodicity) limits every axis to two domains, imcon 6 (slab) the z axis (domains.F90:103-105); where no
    factorisation fits, error 520."""
    cell = np.diag([40.0, 40.0, 40.0]).reshape(9)
    out = run_check("host", dict(cell=cell, dd_cases=np.array([6, 8, 0, 6, 16, 5, 0, 8, 0], dtype=np.int32),
                                 dd_widths=np.array([40.0, 40.0, 40.0] * 3)), tmp_path)
    res = out["dd_results"].reshape(-1, 58)
    assert tuple(res[0, :3]) == (2, 2, 2)
    assert res[1, 2] <= 2 and res[1, 0] * res[1, 1] * res[1, 2] == 16
    assert tuple(res[2, :3]) == (2, 2, 2)
    err = run_check("host", dict(cell=cell, dd_cases=np.array([0, 16, 0], dtype=np.int32), dd_widths=np.array([40.0] * 3)),
                    tmp_path, expect_rc=255)
    assert "520" in err


# ---------------------------------------------------------------- CPU: force-field tables
@pytest.mark.parametrize("name", ["argon_126", "argon_lj_direct_fs", "nacl_ewald", "nacl_rvdw", "water", "nacl_fscp_damped",
                                  "nacl_rfp", "nacl_dddp", "nacl_partial"])
def test_cpp_tables_equal_the_python_host_and_oracle(oracle, tmp_path, name):
    s, spec = {
        "argon_126": lambda: (systems.argon(4), SPEC_AR_126),
        "argon_lj_direct_fs": lambda: (systems.argon(4, form="lj", direct=True, force_shift=True), SPEC_AR_LJ),
        "nacl_ewald": lambda: (systems.nacl(3, rcut=12.0, padding=0.24), SPEC_NACL),
        "nacl_rvdw": lambda: (systems.nacl(3, rcut=8.0, padding=0.2, rvdw=6.5), SPEC_NACL),
        "water": lambda: (systems.spce_water(64, rcut=8.0, padding=0.2), SPEC_WATER),
        "nacl_fscp_damped": lambda: (systems.nacl(3, rcut=8.0, padding=0.2, coulomb="fscp", damping=0.2), SPEC_NACL),
        "nacl_rfp": lambda: (systems.nacl(3, rcut=8.0, padding=0.2, coulomb="rfp", eps=78.0), SPEC_NACL),
        "nacl_dddp": lambda: (systems.nacl(3, rcut=8.0, padding=0.2, coulomb="dddp", eps=2.5), SPEC_NACL),
        "nacl_partial": lambda: (systems.nacl(3, rcut=8.0, padding=0.2, vdw_pairs=((1, 2),)), [SPEC_NACL[1]]),
    }[name]()
    recs = ff_records(s, spec)
    recs["ew_precision"] = [1.0e-6]
    out = run_check("host", recs, tmp_path)
    check_ff_equal(out, s.ff)
    # and straight against the oracle's generators
    g = s.ff.mxgrid
    for k, (ai, aj, form, p) in enumerate(spec):
        tp, tf = oracle.vdw_generate(tables.KEYPOT[form], p, s.ff.rvdw, g)
        assert np.array_equal(out["tab_potential"].reshape(-1, g + 1)[k], tp)
        assert np.array_equal(out["tab_force"].reshape(-1, g + 1)[k], tf)
        if s.ff.force_shift and s.ff.direct:
            assert (out["afs"][k], out["bfs"][k]) == oracle.vdw_direct_fs(tables.KEYPOT[form], p, s.ff.rvdw)
    if s.ff.ew_active or s.ff.coul_damp:
        n = s.ff.ew_n
        e, d, rs = oracle.erfcgen(s.rcut, s.ff.alpha, n)
        assert np.array_equal(out["erfc"], e) and np.array_equal(out["erfc_deriv"], d) and out["electro"][1] == rs
        if s.ff.coul_damp:                                   # coul_spole.F90:192-195 from the oracle's tables
            fs = d[n - 4] * s.rcut
            assert out["electro"][2] == fs and out["electro"][3] == -(e[n - 4] + fs * s.rcut)
    L = oracle.lib()
    assert out["sizes"][0] == L.ora_max_grid(s.rcut)
    assert out["sizes"][1] == L.ora_max_list(s.density, s.rcut + s.padding)
    assert out["alpha_from_precision"][0] == L.ora_ewald_alpha(1.0e-6, s.rcut)


@pytest.mark.parametrize("remake", [False, True])
def test_cpp_vdw_table_read(oracle, tmp_path, remake):
    """TABLE file -> tab_potential / tab_force: same bits as the oracle's restatement of vdw.F90:1196-1341, on the file's own
    grid and re-gridded (delpot != dlrpot); FIELD declares the pairs as 'tab'."""
    rvdw = 8.0
    g = tables.max_grid(rvdw)
    dlrpot = rvdw / (g - 4)
    delpot, ngrid = (0.0075, 1150) if remake else (dlrpot, g)
    r = np.arange(1, ngrid + 1) * delpot
    pairs = []
    for (ai, aj), p in BHM.items():
        e, gm = tables.pot_energy(tables.VDW_BHM, p, r)
        pairs.append((["Na", "Cl"][ai - 1], ["Na", "Cl"][aj - 1], 0.25 * ai, -0.5 * aj, e, gm))
    path = str(tmp_path / "TABLE")
    tables.write_table_file(path, pairs, delpot, ngrid * delpot, ngrid)
    s = systems.nacl(3, rcut=rvdw, padding=0.2)
    recs = ff_records(s, [(ai, aj, "tab", []) for (ai, aj) in BHM])
    recs.update(table_file=path, unique_atom="Na\nCl", engunit=[1.5], force_shift=[1])
    out = run_check("host", recs, tmp_path)
    tp = out["tab_potential"].reshape(-1, g + 1)
    tf = out["tab_force"].reshape(-1, g + 1)
    for k, (a1, a2, elrc, vlrc, e, gm) in enumerate(pairs):
        rp = oracle.vdw_table_regrid(e, delpot, rvdw, g, False, 1.5)
        rf = oracle.vdw_table_regrid(gm, delpot, rvdw, g, True, 1.5)
        rp[g - 3:g - 1] = 0.0                   # vdw.F90:1343-1352 (force-shifted tables end in zeros)
        rf[g - 3:g - 1] = 0.0
        assert np.array_equal(tp[k], rp) and np.array_equal(tf[k], rf)
        assert out["param"].reshape(-1, 7)[k, 0] == elrc * 1.5 and out["param"].reshape(-1, 7)[k, 1] == vlrc * 1.5


def test_cpp_vdw_table_read_errors(tmp_path):
    """The reference's error exits: 24 (end of file), 22 (grid too coarse), 0 (too few points / cutoff), 81 (unknown label)."""
    rvdw = 8.0
    g = tables.max_grid(rvdw)
    s = systems.nacl(3, rcut=rvdw, padding=0.2)
    base = ff_records(s, [(1, 2, "tab", [])])
    base.update(unique_atom="Na\nCl")
    r = np.arange(1, g + 1) * (rvdw / (g - 4))
    e, gm = tables.pot_energy(tables.VDW_BHM, BHM[(1, 2)], r)

    def kode(text_or_pairs, delpot=rvdw / (g - 4), cut=None, ngrid=g, truncate=None):
        path = str(tmp_path / "TABLE_bad")
        tables.write_table_file(path, text_or_pairs, delpot, ngrid * delpot if cut is None else cut, ngrid)
        if truncate is not None:
            lines = open(path).read().split("\n")
            open(path, "w").write("\n".join(lines[:truncate]))
        recs = dict(base)
        recs.update(bad_table_file=path, table_file=path)
        del recs["table_file"]
        return int(run_check("host", recs, tmp_path)["bad_table_kode"][0])

    good = [("Na", "Cl", 0.0, 0.0, e, gm)]
    assert kode(good) == -1                                                      # reads fine
    assert kode(good, truncate=40) == 24                                         # end of file inside the arrays
    assert kode([("Na", "Cl", 0.0, 0.0, e[:900], gm[:900])], ngrid=900) == 0     # fewer points than max_grid - 4
    assert kode([("Na", "Cl", 0.0, 0.0, e[:1100], gm[:1100])], delpot=0.02, ngrid=1100) == 22   # coarser than delr_max
    assert kode(good, cut=rvdw - 1.0) == 0                                       # cutpot < rvdw
    assert kode([("Na", "K", 0.0, 0.0, e, gm)]) == 81                            # label not among the site types
    assert kode([("Na", "Na", 0.0, 0.0, e, gm)]) == 0                            # FIELD says Na-Cl is the tabulated pair


# ---------------------------------------------------------------- CPU: vnl_check decision logic
def _ora_vnl_trace(oracle, l_str, bspline, cutoff, padding, cell, dims, tols):
    import ctypes as C
    L = oracle.lib()
    io = np.array([padding, cutoff + padding])
    flags = np.array([1, 1], dtype=np.int32)
    ns = np.array([0.0, 0.0, 0.0, 999999999.0, 0.0])
    c = np.ascontiguousarray(cell, dtype=np.float64)
    d = np.ascontiguousarray(dims, dtype=np.int32)
    rows, kodes = [], []
    for t in tols:
        width = C.c_double(0.0)
        rc = L.ora_vnl_decide(C.c_int(int(l_str)), C.c_double(t), C.c_int(bspline), C.c_double(cutoff), C.c_void_p(io.ctypes.data),
                              C.c_void_p(c.ctypes.data), C.c_void_p(d.ctypes.data), C.c_void_p(flags.ctypes.data),
                              C.c_void_p(ns.ctypes.data), C.byref(width))
        kodes.append(rc)
        rows.append([float(flags[0]), io[0], io[1], width.value] + list(ns) + [float(flags[1])])
        if rc:
            break
    return np.array(rows), kodes


@pytest.mark.parametrize("l_str,bspline,cell,mxnode,rcut,padding", [
    (True, 1, [98.78, 98.78, 98.78], 1, 12.0, 0.24),       # C2, strict: padding stays what CONTROL said
    (False, 1, [98.78, 98.78, 98.78], 1, 12.0, 0.24),      # no strict + SPME: pushed to 2 % of rcut = 0.24 (no change)
    (False, 0, [114.39, 114.39, 114.39], 1, 8.5, 0.1),     # no strict, no SPME: re-tuned to 4 % of rcut
    (False, 0, [40.0, 40.0, 40.0], 8, 8.0, 0.2),           # two link cells per domain: the m9 slack branch (mxnode > 1)
    (False, 0, [19.0, 19.0, 19.0], 1, 8.0, 0.2),           # serial exception branch (0.5 width - rcut)
    (False, 0, [17.2, 17.2, 17.2], 8, 8.0, 0.7),           # no link cell fits rcut + padding: padding reset with slack
    (True, 0, [17.2, 17.2, 17.2], 8, 8.0, 0.7),            # the same in strict mode: error 307
    (False, 0, [15.0, 15.0, 15.0], 8, 8.0, 0.2),           # domain narrower than rcut: error 307
])
def test_cpp_vnl_check_decisions_equal_the_oracle(oracle, tmp_path, l_str, bspline, cell, mxnode, rcut, padding):
    """neighbours.F90:182-284: update test, the padding re-tune of the 'no strict' regime, error 307, skip statistics."""
    cellm = np.diag(cell).reshape(9)
    rng = np.random.default_rng(3)
    tols = list(rng.uniform(0.0, 0.2, 40))
    recs = dict(cell=cellm, imcon=[1], megatm=[1000], rcut=[rcut], padding=[padding], mxnode=[mxnode], idnode=[0],
                vnl_tols=np.array(tols), l_str=[int(l_str)], bspline=[bspline])
    out = run_check("host", recs, tmp_path)
    w = oracle.World(mxnode, cellm, 1)
    dims = w.dd(0)[0][:3]
    rows, kodes = _ora_vnl_trace(oracle, l_str, bspline, rcut, padding, cellm, dims, tols)
    assert list(out["vnl_kode"]) == kodes
    got = out["vnl_trace"].reshape(-1, 10)
    if kodes[-1] == 0:
        assert np.array_equal(got, rows)
        assert got[:, 0].min() == 0.0 and got[:, 0].max() == 1.0      # both outcomes occurred
    else:
        assert kodes[-1] == 307


# ---------------------------------------------------------------- GPU: the drop-in call sequence through the C++ host
def _dropin_records(s, spec, w, rank, P):
    d = domain_inputs(w, rank)
    parts = d["parts"].copy()
    for k in ("fxx", "fyy", "fzz"):
        parts[k] = 0.0
    recs = ff_records(s, spec)
    recs.update(mxnode=[P], idnode=[rank], natms=[d["natms"]], nlast=[d["nlast"]], parts=parts, ltype=d["ltype"], ltg=d["ltg"],
                lfrzn=d["lfrzn"], lbook=[int(s.lbook)], max_list=[d["max_list"]], force_mode=[1])
    if s.lbook:
        recs.update(max_exclude=[d["list_excl"].shape[1] - 1], list_excl=d["list_excl"])
    return d, recs


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["argon", "nacl", "water", "nacl_8_domains", "nacl_rfp"])
def test_cpp_host_dropin_against_oracle(tmp_path, name):
    s, spec, P, rank = {
        "argon": lambda: (systems.argon(6), SPEC_AR_126, 1, 0),
        "nacl": lambda: (systems.nacl(4, rcut=8.0, padding=0.2), SPEC_NACL, 1, 0),
        "water": lambda: (systems.spce_water(512, rcut=8.0, padding=0.2), SPEC_WATER, 1, 0),
        "nacl_8_domains": lambda: (systems.nacl(8, rcut=8.0, padding=0.2), SPEC_NACL, 8, 5),
        "nacl_rfp": lambda: (systems.nacl(4, rcut=8.0, padding=0.2, coulomb="rfp", eps=5.0, damping=0.25), SPEC_NACL, 1, 0),
    }[name]()
    w = world_for(s, P=P)
    w.two_body()
    d, recs = _dropin_records(s, spec, w, rank, P)
    recs.update(vnl_shift=[0.06, 0.0, 0.08], l_str=[1], bspline=[1])
    out = run_check("dropin", recs, tmp_path)
    ref = w.list(rank)
    lst = out["list"].reshape(ref.shape)
    assert np.array_equal(lst[:, :4], ref[:, :4])                                  # bit-exact counters, members and order
    used = np.arange(ref.shape[1] - 4)[None, :] < ref[:, 1:2]
    assert np.array_equal(np.where(used, lst[:, 4:], 0), np.where(used, ref[:, 4:], 0))
    parts = np.frombuffer(out["parts"], dtype=COREPART)
    a, b = force_errors(parts_forces(parts, d["natms"]), parts_forces(d["parts"], d["natms"]))
    assert a <= FORCE_TOL and b <= 1.0e-7
    assert per_atom_force_error(parts_forces(parts, d["natms"]), parts_forces(d["parts"], d["natms"]))["per_atom_significant"] <= FORCE_TOL
    oo = w.results(rank)
    scale = max(abs(oo[:6]).max(), 1.0)
    for k in range(6):
        assert abs(out["sums"][k] - oo[k]) <= ENERGY_TOL * max(abs(oo[k]), 1e-6 * scale), (k, out["sums"][k], oo[k])
    assert np.abs(out["stress"] - oo[6:15]).max() <= ENERGY_TOL * abs(oo[6:15]).max()
    # vnl_check after a rigid 0.1 A shift of the local atoms: tol = 0.1 >= half_minus * padding iff padding <= 0.2
    upd, padding, rx = out["vnl"][:3]
    assert upd == (1.0 if 0.1 >= np.nextafter(0.5, 0.0) * s.padding else 0.0)
    assert padding == s.padding and rx == s.rcut + s.padding
    assert out["launches"][0] > 0


@pytest.mark.gpu
def test_cpp_host_native_md_against_oracle(tmp_path):
    """md_vv around the path, device-resident, driven from C++: the same rebuild decisions and energies as the oracle's
    trajectory (vv stage 1, vnl_check, relocate + halo + list or refresh, two_body_forces, vv stage 2)."""
    s = systems.nacl(4, rcut=8.0, padding=0.3, temperature=1200.0)
    nsteps, dt = 25, 0.002
    recs = ff_records(s, SPEC_NACL)
    recs.update(type_site=s.type_site, charge_site=s.charge_site, freeze_site=s.freeze_site, weight_site=s.weight_site,
                xyz=s.xyz, vel=s.vel, ltg=np.arange(1, s.megatm + 1, dtype=np.int32), lsite=s.lsite, nsteps=[nsteps],
                timestep=[dt], force_mode=[1])
    from oracle import oracle as ora
    w = ora.World.from_system(s, P=1)
    # the C++ driver folds the CONFIG positions itself (read_config_fold), like the oracle's load
    out = run_check("md", recs, tmp_path)
    sums = out["sums"].reshape(nsteps + 1, 16)
    w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
    oo = w.two_body()
    rebuilt = []
    for step in range(nsteps + 1):
        if step > 0:
            w.vv(1, dt, s.weight_by_type)
            upd, tol = w.vnl_check()
            if upd:
                w.relocate(); w.set_halo(); assert w.link_cell_pairs() == 0
                rebuilt.append(step)
            else:
                assert w.refresh_halo() == 0
            oo = w.two_body()
            w.vv(2, dt, s.weight_by_type)
        tot_o, tot_g = oo[0] + oo[2], sums[step, 0] + sums[step, 2]
        assert abs(tot_g - tot_o) <= 1e-8 * abs(tot_o), (step, sums[step, :4], oo[:4])
    assert list(out["rebuilt_at"]) == rebuilt and len(rebuilt) >= 1
    c = w.counts(0)
    assert tuple(out["counts"]) == (c["natms"], c["nlast"])
    parts = np.frombuffer(out["parts"], dtype=COREPART)
    po = w.parts(0)
    order_g, order_o = np.argsort(out["ltg"][:c["natms"]]), np.argsort(w.ints(0)["ltg"][:c["natms"]])
    for k in ("xxx", "yyy", "zzz"):
        assert np.abs(parts[k][:c["natms"]][order_g] - po[k][:c["natms"]][order_o]).max() < 1e-7


# ---------------------------------------------------------------- CPU: long-range corrections and the end of two_body_forces
def _ora_lrc(oracle, ff, num_type, numfrz, imcon, volm):
    import ctypes as C
    L = oracle.lib()
    lst, ltp = np.ascontiguousarray(ff.vdw_list_c, dtype=np.int32), np.ascontiguousarray(ff.ltp, dtype=np.int32)
    par = np.ascontiguousarray(ff.param, dtype=np.float64)
    nt, nf = np.ascontiguousarray(num_type, dtype=np.float64), np.ascontiguousarray(numfrz, dtype=np.float64)
    e, v = C.c_double(), C.c_double()
    vp = lambda a: C.c_void_p(a.ctypes.data)
    L.ora_vdw_lrc(C.c_int(ff.ntypes), vp(lst), vp(ltp), vp(par), C.c_double(ff.rvdw), C.c_int(int(ff.force_shift)), C.c_int(imcon),
                  C.c_double(volm), vp(nt), vp(nf), C.byref(e), C.byref(v))
    return e.value, v.value


def _ora_epilogue(oracle, sums8, spme, sumchg, alpha, eps, volm, elrc, vlrc, mxnode, stress_in):
    import ctypes as C
    L = oracle.lib()
    in6 = np.ascontiguousarray(sums8[:6], dtype=np.float64)
    tot, st = np.zeros(4), np.ascontiguousarray(stress_in, dtype=np.float64).copy()
    vp = lambda a: C.c_void_p(a.ctypes.data)
    L.ora_two_body_epilogue(vp(in6), C.c_double(sums8[6]), C.c_double(sums8[7]), C.c_int(int(spme)), C.c_double(sumchg), C.c_double(alpha),
                            C.c_double(eps), C.c_double(volm), C.c_double(elrc), C.c_double(vlrc), C.c_int(mxnode), vp(tot), vp(st))
    return tot, st


@pytest.mark.parametrize("name", ["argon_126", "argon_lj", "nacl_bhm", "buck", "argon_shifted", "slab"])
def test_vdw_lrc_and_two_body_totals(oracle, tmp_path, name):
    """vdw_lrc (vdw.F90:617-967) and the end of two_body_forces (two_body.F90:672-790): the oracle's restatement is pinned by
    numerical integration of the potential beyond the cutoff (elrc = 2 pi N_i N_j / V * int u r^2 dr, vlrc from r du/dr); the
    C++ and Python hosts equal the oracle bit for bit."""
    from scipy.integrate import quad
    s, spec = {
        "argon_126": lambda: (systems.argon(4), SPEC_AR_126),
        "argon_lj": lambda: (systems.argon(4, form="lj"), SPEC_AR_LJ),
        "nacl_bhm": lambda: (systems.nacl(3, rcut=9.0, padding=0.2), SPEC_NACL),
        "buck": lambda: (systems.nacl(3, rcut=9.0, padding=0.2, vdw_pairs=()), [(1, 2, "buck", [1.0e5, 0.31, 650.0]), (2, 2, "buck", [2.2e5, 0.29, 2800.0])]),
        "argon_shifted": lambda: (systems.argon(4, form="lj", force_shift=True), SPEC_AR_LJ),
        "slab": lambda: (systems.argon(4), SPEC_AR_126),
    }[name]()
    if name == "buck":
        ff = tables.ForceField(2, 9.0, 9.0)
        for ai, aj, form, p in spec:
            ff.add(ai, aj, form, p)
        ff.set_ewald(precision=1e-6)
        s.ff = ff.finalize()
    imcon = 6 if name == "slab" else s.imcon
    types = s.type_site[s.lsite - 1]
    num_type = np.bincount(types, minlength=s.ff.ntypes + 1)[1:].astype(np.float64)
    numfrz = np.zeros_like(num_type)
    if name == "nacl_bhm":
        numfrz[:] = [3.0, 5.0]
    sumchg = 0.0 if name != "buck" else 2.0                     # a net charge switches Fuchs' correction on
    e_o, v_o = _ora_lrc(oracle, s.ff, num_type, numfrz, imcon, s.volume)
    # second opinion: numerical integration of every pair's analytic form beyond rvdw
    e_q = v_q = 0.0
    if not s.ff.force_shift and imcon not in (0, 6):
        for ai, aj, form, p in spec:
            pp = np.zeros(7); pp[:len(p)] = p
            u = lambda r: float(tables.pot_energy(tables.KEYPOT[form], pp, np.array([r]))[0][0]) * r * r
            g = lambda r: float(tables.pot_energy(tables.KEYPOT[form], pp, np.array([r]))[1][0]) * r * r
            mult = 1.0 if ai == aj else 2.0
            # the reference drops the exponential tails of buck / bhm: integrate the dispersion part only
            if form in ("buck", "bhm"):
                c6, d8 = (pp[2], 0.0) if form == "buck" else (pp[3], pp[4])
                u = lambda r, c6=c6, d8=d8: (-c6 / r ** 6 - d8 / r ** 8) * r * r
                g = lambda r, c6=c6, d8=d8: (-6.0 * c6 / r ** 6 - 8.0 * d8 / r ** 8) * r * r
            dens = 2.0 * np.pi * (num_type[ai - 1] * num_type[aj - 1] - numfrz[ai - 1] * numfrz[aj - 1]) / s.volume ** 2
            e_q += s.volume * dens * mult * quad(u, s.ff.rvdw, np.inf, epsabs=0, epsrel=1e-12)[0]
            v_q += -s.volume * dens * mult * quad(g, s.ff.rvdw, np.inf, epsabs=0, epsrel=1e-12)[0]
    assert abs(e_o - e_q) <= 1e-9 * max(abs(e_q), 1e-300) and abs(v_o - v_q) <= 1e-9 * max(abs(v_q), 1e-300)
    if name in ("argon_shifted", "slab"):
        assert (e_o, v_o) == (0.0, 0.0)
    else:
        assert e_o < 0.0 and v_o > 0.0                            # attractive dispersion tail beyond the cutoff (virial = -r dU/dr)
    assert tables.vdw_lrc(s.ff, num_type, numfrz, imcon, s.volume) == (e_o, v_o)
    # the end of two_body_forces on made-up partial sums
    rng = np.random.default_rng(17)
    sums8 = rng.normal(0.0, 1.0e5, 8)
    stress_in = rng.normal(0.0, 1.0e4, 9)
    mxnode = 8
    spme = s.ff.ew_active
    tot_o, st_o = _ora_epilogue(oracle, sums8, spme, sumchg, s.ff.alpha, s.ff.eps, s.volume, e_o, v_o, mxnode, stress_in)
    recs = ff_records(s, spec)
    recs.update(imcon=[imcon], num_type=num_type, numfrz=numfrz, volm=[s.volume], sumchg=[sumchg], partial_sums=sums8, stress_in=stress_in,
                mxnode=[mxnode])
    out = run_check("host", recs, tmp_path)
    assert tuple(out["lrc"]) == (e_o, v_o)
    assert np.array_equal(out["totals"], tot_o) and np.array_equal(out["stress_out"], st_o)
    py = tables.two_body_totals(np.concatenate([sums8[:6], stress_in, [0.0]]), e_o, v_o, mxnode, spme, sumchg, s.ff.alpha, s.ff.eps, s.volume,
                                sums8[6], sums8[7])
    assert np.array_equal(np.array(py[:4]), tot_o) and np.array_equal(py[4], st_o)
    if name == "buck":
        assert tot_o[0] != sums8[6] + sums8[2] + sums8[4]         # the net-charge term is in


def test_cpp_host_logic_fuzz_against_the_oracle(oracle, tmp_path):
    """Random cells, decompositions, cutoffs and displacement sequences: map_domains and every branch of the vnl_check
    decision logic (incl. the error-307 exits) give the oracle's answers bit for bit."""
    rng = np.random.default_rng(2026)
    seen_kodes, seen_retune = set(), 0
    for case in range(60):
        if case % 3 == 0:
            cellm = np.diag(rng.uniform(14.0, 120.0, 3)).reshape(9)
        elif case % 3 == 1:
            L = rng.uniform(14.0, 120.0)
            cellm = np.diag([L, L, L]).reshape(9)
        else:
            cellm = (np.diag(rng.uniform(20.0, 90.0, 3)) + rng.uniform(-3.0, 3.0, (3, 3))).reshape(9)
        P = int(rng.choice([1, 2, 3, 4, 6, 8, 12, 16, 24, 27, 32, 64]))
        rcut = float(rng.uniform(5.0, 14.0))
        padding = float(rng.choice([0.05, 0.1, 0.2, 0.35, 0.6, 1.0]))
        l_str, bspline = bool(rng.integers(0, 2)), int(rng.integers(0, 2)) * 8
        tols = list(rng.uniform(0.0, 0.6 * padding + 0.05, 12))
        w = oracle.World(P, cellm, 1)
        dd6, map26 = w.dd(P - 1)
        recs = dict(cell=cellm, imcon=[1], megatm=[1000], rcut=[rcut], padding=[padding], mxnode=[P], idnode=[0],
                    vnl_tols=np.array(tols), l_str=[int(l_str)], bspline=[bspline],
                    dd_cases=np.array([1, P, P - 1], dtype=np.int32), dd_widths=oracle.dcell(cellm)[6:9])
        out = run_check("host", recs, tmp_path)
        res = out["dd_results"]
        assert np.array_equal(res[:6], dd6) and np.array_equal(res[6:32], map26), (case, P)
        rows, kodes = _ora_vnl_trace(oracle, l_str, bspline, rcut, padding, cellm, dd6[:3], tols)
        assert list(out["vnl_kode"]) == kodes, (case, kodes)
        got = out["vnl_trace"].reshape(-1, 10)
        if kodes[-1] == 0:
            assert np.array_equal(got, rows), case
            seen_retune += int(len(set(rows[:, 1])) > 1 or rows[0, 1] != padding)
        seen_kodes.add(kodes[-1])
    assert seen_kodes == {0, 307} and seen_retune >= 5


def test_cpp_tables_fuzz_against_the_oracle(oracle, tmp_path):
    """Random parameters for the four analytic forms, random cutoffs (table sizes), force-shift constants, random Ewald
    alpha: vdw_generate / vdw_direct_fs_generate / erfcgen of the C++ host equal the oracle's bit for bit."""
    rng = np.random.default_rng(77)
    forms = {"12-6": lambda: [rng.uniform(1e5, 1e8), rng.uniform(1e2, 1e4)], "lj": lambda: [rng.uniform(10, 500), rng.uniform(2.0, 4.0)],
             "buck": lambda: [rng.uniform(1e4, 1e6), rng.uniform(0.2, 0.4), rng.uniform(10, 5e3)],
             "bhm": lambda: [rng.uniform(1e3, 5e3), rng.uniform(2.5, 3.5), rng.uniform(2.0, 3.5), rng.uniform(1e4, 1e6), rng.uniform(1e3, 1e6)]}
    for case in range(12):
        rvdw = float(rng.uniform(6.0, 14.0))
        rcut = rvdw + float(rng.choice([0.0, 0.5, 1.5]))
        alpha = float(rng.uniform(0.15, 0.45))
        names = list(forms)
        spec = []
        for k, (ai, aj) in enumerate([(1, 1), (1, 2), (2, 2)]):
            f = names[int(rng.integers(0, 4))]
            spec.append((ai, aj, f, forms[f]()))
        pairs, prm = [], []
        for ai, aj, f, p in spec:
            pairs += [ai, aj, tables.KEYPOT[f]]
            q = np.zeros(7); q[:len(p)] = p
            prm.append(q)
        recs = dict(cell=np.diag([60.0, 60.0, 60.0]).reshape(9), ntypes=[2], rvdw=[rvdw], rcut=[rcut], padding=[0.2], force_shift=[1], direct=[1],
                    pot_pairs=np.array(pairs, dtype=np.int32), pot_param=np.array(prm).reshape(-1), electro_key=[1], eps=[1.0],
                    damping=[0.0], fdens=[0.03], ew_alpha=[alpha])
        out = run_check("host", recs, tmp_path)
        g = oracle.lib().ora_max_grid(rvdw)
        assert tuple(out["vdw_sizes"]) == (3, 3, g)
        for k, (ai, aj, f, p) in enumerate(spec):
            tp, tf = oracle.vdw_generate(tables.KEYPOT[f], p, rvdw, g)
            assert np.array_equal(out["tab_potential"].reshape(-1, g + 1)[k], tp), (case, f)
            assert np.array_equal(out["tab_force"].reshape(-1, g + 1)[k], tf), (case, f)
            assert (out["afs"][k], out["bfs"][k]) == oracle.vdw_direct_fs(tables.KEYPOT[f], p, rvdw), (case, f)
        n = oracle.lib().ora_max_grid(rcut)
        e, d, rs = oracle.erfcgen(rcut, alpha, n)
        assert np.array_equal(out["erfc"], e) and np.array_equal(out["erfc_deriv"], d) and out["electro"][1] == rs
