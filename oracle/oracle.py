"""ctypes front-end of the CPU oracle (oracle/dlp_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs -- never from
the product package.  ``build()`` compiles the restatement with the recipe in oracle/Makefile.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libdlp_oracle.so")
PERF_PATH = os.path.join(HERE, "_build", "libdlp_oracle_perf.so")

COREPART = np.dtype([("xxx", "f8"), ("yyy", "f8"), ("zzz", "f8"), ("fxx", "f8"), ("fyy", "f8"), ("fzz", "f8"),
                     ("chge", "f8"), ("pad1", "i4"), ("pad2", "i4")])
assert COREPART.itemsize == 64


def _host_fingerprint():
    """Identifies the CPU the perf variant (-march=native) was built for: it must never run on another machine's cores (the
    built file travels with the repo snapshot to the GPU box)."""
    import hashlib
    import platform
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        flags = [ln.split(":", 1)[1] for ln in txt.split("\n") if ln.startswith(("flags", "model name"))][:2]
    except OSError:
        flags = []
    return hashlib.sha1((platform.machine() + "|" + "|".join(flags)).encode()).hexdigest()


def build(perf=False, quiet=True):
    target = "perf" if perf else "all"
    src = os.path.join(HERE, "dlp_oracle.cpp")
    out = PERF_PATH if perf else LIB_PATH
    stamp = out + ".host"
    fresh = os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src)
    if perf and fresh:
        try:
            fresh = open(stamp).read().strip() == _host_fingerprint()
        except OSError:
            fresh = False
    if fresh:
        return out
    if perf and os.path.exists(out):
        os.remove(out)                      # make would consider it up to date
    subprocess.run(["make", "-C", HERE, target], check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.STDOUT if quiet else None)
    if perf:
        with open(stamp, "w") as f:
            f.write(_host_fingerprint())
    return out


_libs = {}


def _vp(a):
    return C.c_void_p(a.ctypes.data)


def lib(perf=False):
    key = bool(perf)
    if key in _libs:
        return _libs[key]
    path = PERF_PATH if perf else LIB_PATH
    if not os.path.exists(path):
        build(perf=perf)
    L = C.CDLL(path)
    L.ora_world_create.restype = C.c_void_p
    L.ora_world_create.argtypes = [C.c_int, C.c_void_p, C.c_int]
    L.ora_brute_pairs.restype = C.c_long
    L.ora_calc_erfc.restype = C.c_double
    L.ora_calc_erfc.argtypes = [C.c_double]
    L.ora_ewald_alpha.restype = C.c_double
    L.ora_ewald_alpha.argtypes = [C.c_double, C.c_double]
    L.ora_max_grid.argtypes = [C.c_double]
    L.ora_max_list.argtypes = [C.c_double, C.c_double]
    _libs[key] = L
    return L


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---- stateless helpers ------------------------------------------------------------------------------------------
def dcell(cell):
    out = np.zeros(10)
    c = _d(cell)
    lib().ora_dcell(_vp(c), _vp(out))
    return out


def invert(cell):
    b = np.zeros(9)
    det = C.c_double()
    c = _d(cell)
    lib().ora_invert(_vp(c), _vp(b), C.byref(det))
    return b, det.value


def match(n, lst):
    l = _i(lst)
    return bool(lib().ora_match(C.c_int(n), C.c_int(len(l)), _vp(l)))


def images(imcon, cell, x, y, z):
    x, y, z = _d(x).copy(), _d(y).copy(), _d(z).copy()
    c = _d(cell)
    lib().ora_images(C.c_int(imcon), _vp(c), C.c_int(len(x)), _vp(x), _vp(y), _vp(z))
    return x, y, z


def pot_energy(keypot, param, r):
    p = np.zeros(7)
    p[:len(param)] = param
    e, g = C.c_double(), C.c_double()
    lib().ora_pot_energy(C.c_int(keypot), _vp(p), C.c_double(r), C.byref(e), C.byref(g))
    return e.value, g.value


def kat_vdw_direct(keypot, param):
    """Replays source/unit_tests/test_vdw.F90:42-177 for one potential -> (energy, virial)."""
    p = np.zeros(7)
    p[:len(param)] = param
    e, v = C.c_double(), C.c_double()
    lib().ora_kat_vdw_direct(C.c_int(keypot), _vp(p), C.byref(e), C.byref(v))
    return e.value, v.value


def erfcgen(rcut, alpha, nsamples):
    et = np.zeros(nsamples + 1)
    dt = np.zeros(nsamples + 1)
    rs = C.c_double()
    lib().ora_erfcgen(C.c_double(rcut), C.c_double(alpha), C.c_int(nsamples), _vp(et), _vp(dt), C.byref(rs))
    return et, dt, rs.value


def vdw_generate(keypot, param, rvdw, mxgrid):
    p = np.zeros(7)
    p[:len(param)] = param
    tp = np.zeros(mxgrid + 1)
    tf = np.zeros(mxgrid + 1)
    lib().ora_vdw_generate(C.c_int(keypot), _vp(p), C.c_double(rvdw), C.c_int(mxgrid), _vp(tp), _vp(tf))
    return tp, tf


def vdw_direct_fs(keypot, param, rvdw):
    p = np.zeros(7)
    p[:len(param)] = param
    a, b = C.c_double(), C.c_double()
    lib().ora_vdw_direct_fs(C.c_int(keypot), _vp(p), C.c_double(rvdw), C.byref(a), C.byref(b))
    return a.value, b.value


def vdw_table_regrid(buf, delpot, rvdw, mxgrid, is_force, engunit=1.0):
    b = _d(buf)
    tab = np.zeros(mxgrid + 1)
    lib().ora_vdw_table_regrid(_vp(b), C.c_int(len(b)), C.c_double(delpot), C.c_double(rvdw), C.c_int(mxgrid),
                               C.c_int(int(is_force)), C.c_double(engunit), _vp(tab))
    return tab


def brute_pairs(xyz, cell, rx):
    x = _d(xyz)
    c = _d(cell)
    n = x.shape[0]
    band = C.c_long()
    L = lib()
    cnt = L.ora_brute_pairs(C.c_int(n), _vp(x), _vp(c), C.c_double(rx), None, C.c_long(0), C.byref(band))
    out = np.zeros((cnt, 2), dtype=np.int32)
    L.ora_brute_pairs(C.c_int(n), _vp(x), _vp(c), C.c_double(rx), _vp(out), C.c_long(cnt), C.byref(band))
    return out, band.value


# ---- world ------------------------------------------------------------------------------------------------------
class World:
    """In-process multi-domain restatement of the reference's short-range engine (one Dom per MPI rank)."""

    def __init__(self, P, cell, imcon=1, perf=False):
        self.L = lib(perf)
        self.P = P
        c = _d(cell)
        self.cell = c.copy()
        self.h = C.c_void_p(self.L.ora_world_create(C.c_int(P), _vp(c), C.c_int(imcon)))

    def __del__(self):
        try:
            self.L.ora_world_destroy(self.h)
        except Exception:
            pass

    @classmethod
    def from_system(cls, sysm, P=1, fold=True, perf=False, ecw=None):
        w = cls(P, sysm.cell, sysm.imcon, perf=perf)
        w.set_cutoffs(sysm.rcut, sysm.padding, sysm.pdplnc, ecw)
        w.set_sites(sysm.type_site, sysm.charge_site, sysm.freeze_site)
        w.set_forcefield(sysm.ff)
        if sysm.excl is not None:
            w.set_excl(sysm.excl)
        w.set_max_list(sysm.max_list)
        w.load(sysm.xyz, sysm.vel, sysm.lsite, fold=fold)
        return w

    def dd(self, rank=0):
        out = np.zeros(6, dtype=np.int32)
        m = np.zeros(26, dtype=np.int32)
        self.L.ora_world_dd(self.h, C.c_int(rank), _vp(out), _vp(m))
        return out, m

    def set_cutoffs(self, rcut, padding, pdplnc=50.0, ecw=None):
        e = _d(ecw) if ecw is not None else None
        self.L.ora_world_set_cutoffs(self.h, C.c_double(rcut), C.c_double(padding), C.c_double(pdplnc),
                                     _vp(e) if e is not None else None)

    def set_sites(self, type_site, charge_site, freeze_site):
        t, q, f = _i(type_site), _d(charge_site), _i(freeze_site)
        self.L.ora_world_set_sites(self.h, C.c_int(len(t)), _vp(t), _vp(q), _vp(f))

    def set_forcefield(self, ff):
        tp, tf = _d(ff.tab_potential), _d(ff.tab_force)
        par, afs, bfs = _d(ff.param), _d(ff.afs), _d(ff.bfs)
        lst, ltp = _i(ff.vdw_list_c), _i(ff.ltp)
        self.L.ora_world_set_vdw(self.h, C.c_int(ff.ntypes), _vp(lst), C.c_int(ff.max_vdw), C.c_int(ff.n_vdw),
                                 _vp(ltp), C.c_int(ff.mxgrid), _vp(tp), _vp(tf), C.c_double(ff.rvdw),
                                 C.c_int(int(ff.force_shift)), C.c_int(int(ff.direct)), _vp(par), _vp(afs),
                                 _vp(bfs))
        if ff.ew_active:
            e, d = _d(ff.erfc), _d(ff.erfc_deriv)
            self.L.ora_world_set_ewald(self.h, C.c_int(1), C.c_double(ff.alpha), C.c_double(ff.scaling), C.c_int(ff.ew_n),
                                       _vp(e), _vp(d), C.c_double(ff.ew_recip))
        elif getattr(ff, "coul_kind", 0):
            rf = _d(ff.coul_rf)
            if ff.coul_damp:
                e, d = _d(ff.erfc), _d(ff.erfc_deriv)
                self.L.ora_world_set_coulomb(self.h, C.c_int(ff.coul_kind), C.c_int(1), C.c_double(ff.scaling),
                                             C.c_double(ff.coul_force_shift), C.c_double(ff.coul_energy_shift), _vp(rf),
                                             C.c_int(ff.ew_n), _vp(e), _vp(d), C.c_double(ff.ew_recip))
            else:
                self.L.ora_world_set_coulomb(self.h, C.c_int(ff.coul_kind), C.c_int(0), C.c_double(ff.scaling),
                                             C.c_double(ff.coul_force_shift), C.c_double(ff.coul_energy_shift), _vp(rf),
                                             C.c_int(0), None, None, C.c_double(0.0))

    def set_excl(self, excl):
        e = _i(excl)
        self.L.ora_world_set_excl(self.h, C.c_int(e.shape[1] - 1), _vp(e), C.c_int(e.shape[0]))

    def set_max_list(self, ml):
        self.L.ora_world_set_max_list(self.h, C.c_int(ml))

    def load(self, xyz, vel, lsite, fold=True):
        x, s = _d(xyz), _i(lsite)
        v = _d(vel) if vel is not None else None
        return self.L.ora_world_load(self.h, C.c_int(x.shape[0]), _vp(x), _vp(v) if v is not None else None,
                                     _vp(s), C.c_int(int(fold)))

    def set_threads(self, n):
        """Host threads for the world-level loops over domains (halo, migration, vnl_check, integrator)."""
        self.L.ora_world_set_threads(self.h, C.c_int(int(n)))

    def relocate(self):
        return self.L.ora_world_relocate(self.h)

    def set_halo(self):
        return self.L.ora_world_set_halo(self.h)

    def refresh_halo(self):
        return self.L.ora_world_refresh_halo(self.h)

    def vnl_check(self):
        tol = C.c_double()
        upd = self.L.ora_world_vnl_check(self.h, C.byref(tol))
        return bool(upd), tol.value

    def neighskip(self):
        out = np.zeros(5)
        self.L.ora_world_neighskip(self.h, _vp(out))
        return out

    def link_cell_pairs(self, nthreads=1):
        return self.L.ora_world_link_cell_pairs(self.h, C.c_int(nthreads))

    def two_body(self, nthreads=1, zero_forces=True):
        out = np.zeros(15)
        self.L.ora_world_two_body(self.h, C.c_int(nthreads), C.c_int(int(zero_forces)), _vp(out))
        return out

    def rdf_collect(self, rdf_list, n_pairs, max_grid):
        """rdf_collect + rdf_excl_collect over every domain; returns the counts as (n_pairs, max_grid)."""
        lst = _i(rdf_list)
        rdf = np.zeros((n_pairs, max_grid))
        self.L.ora_world_rdf_collect(self.h, _vp(lst), C.c_int(n_pairs), C.c_int(max_grid), _vp(rdf))
        return rdf

    def vv(self, stage, dt, weight_by_type):
        wt = _d(weight_by_type)
        self.L.ora_world_vv(self.h, C.c_int(stage), C.c_double(dt), _vp(wt))

    # per-domain accessors
    def counts(self, rank=0):
        out = np.zeros(11, dtype=np.int32)
        self.L.ora_dom_counts(self.h, C.c_int(rank), _vp(out))
        keys = ["natms", "nlast", "max_list", "max_exclude", "nlx", "nly", "nlz", "nlp", "ncells", "nsbcll", "ibig"]
        return dict(zip(keys, (int(v) for v in out)))

    def parts(self, rank=0):
        n = self.counts(rank)["nlast"]
        p = np.zeros(n, dtype=COREPART)
        self.L.ora_dom_get_parts(self.h, C.c_int(rank), _vp(p))
        return p

    def set_parts(self, rank, parts):
        p = np.ascontiguousarray(parts)
        self.L.ora_dom_set_parts(self.h, C.c_int(rank), _vp(p), C.c_int(len(p)))

    def ints(self, rank=0):
        n = self.counts(rank)["nlast"]
        arrs = [np.zeros(n, dtype=np.int32) for _ in range(5)]
        self.L.ora_dom_get_ints(self.h, C.c_int(rank), *[_vp(a) for a in arrs])
        return dict(zip(["ltg", "lsite", "ltype", "lfrzn", "ixyz"], arrs))

    def vel(self, rank=0):
        n = self.counts(rank)["natms"]
        v = np.zeros((n, 3))
        self.L.ora_dom_get_vel(self.h, C.c_int(rank), _vp(v))
        return v

    def list(self, rank=0):
        c = self.counts(rank)
        out = np.zeros((c["natms"], c["max_list"] + 4), dtype=np.int32)
        self.L.ora_dom_get_list(self.h, C.c_int(rank), _vp(out))
        return out

    def list_excl(self, rank=0):
        c = self.counts(rank)
        out = np.zeros((c["natms"], c["max_exclude"] + 1), dtype=np.int32)
        self.L.ora_dom_get_list_excl(self.h, C.c_int(rank), _vp(out))
        return out

    def cells(self, rank=0):
        c = self.counts(rank)
        wc = np.zeros(c["nlast"], dtype=np.int32)
        al = np.zeros(c["nlast"], dtype=np.int32)
        ls = np.zeros(c["ncells"] + 2, dtype=np.int32)
        self.L.ora_dom_get_cells(self.h, C.c_int(rank), _vp(wc), _vp(al), _vp(ls))
        return wc, al, ls

    def set_collect_pp(self, on=True):
        """stats%collect_pp: the next two_body calls also fill pp_energy / pp_stress (vdw.F90:1741-1755, :1987-2001,
        ewald_spole.F90:205-215)."""
        self.L.ora_world_set_collect_pp(self.h, C.c_int(int(on)))

    def pp(self, rank=0):
        """(pp_energy(natms), pp_stress(natms, 9)) of the last two_body call."""
        n = self.counts(rank)["natms"]
        e, st = np.zeros(n), np.zeros((n, 9))
        assert self.L.ora_dom_get_pp(self.h, C.c_int(rank), _vp(e), _vp(st)) == 0
        return e, st

    def results(self, rank=0):
        out = np.zeros(15)
        self.L.ora_dom_get_results(self.h, C.c_int(rank), _vp(out))
        return out

    def bg(self, rank=0):
        n = self.counts(rank)["nlast"]
        a = [np.zeros(n) for _ in range(3)]
        self.L.ora_dom_get_bg(self.h, C.c_int(rank), *[_vp(x) for x in a])
        return a

    def brute_forces(self, xyz, lsite):
        x, s = _d(xyz), _i(lsite)
        f = np.zeros_like(x)
        out = np.zeros(6)
        self.L.ora_world_brute_forces(self.h, C.c_int(x.shape[0]), _vp(x), _vp(s), _vp(f), _vp(out))
        return f, out

    def gather_forces(self):
        """Forces of all local atoms of all domains ordered by global id."""
        tot = sum(self.counts(r)["natms"] for r in range(self.P))
        f = np.zeros((tot, 3))
        for r in range(self.P):
            n = self.counts(r)["natms"]
            p = self.parts(r)[:n]
            g = self.ints(r)["ltg"][:n] - 1
            f[g, 0], f[g, 1], f[g, 2] = p["fxx"], p["fyy"], p["fzz"]
        return f

    def gather_positions(self):
        tot = sum(self.counts(r)["natms"] for r in range(self.P))
        x = np.zeros((tot, 3))
        for r in range(self.P):
            n = self.counts(r)["natms"]
            p = self.parts(r)[:n]
            g = self.ints(r)["ltg"][:n] - 1
            x[g, 0], x[g, 1], x[g, 2] = p["xxx"], p["yyy"], p["zzz"]
        return x
