"""Host-side mirror of the reference call sites of the short-range path, on top of the libdlpgpu C ABI.

:class:`ShortRange` carries the state the reference keeps in ``neighbours_type`` / ``vdw_type`` /
``electrostatic_type`` for this path and exposes

* ``link_cell_pairs``   -- neighbours.F90:356   (call site drivers.F90:675-679)
* ``two_body_forces``   -- two_body.F90:339-606 (the two pair loops; SPME / gsum / LRC stay with the caller)
* ``vnl_check``         -- neighbours.F90:123   (max displacement; the caller does gmax and the comparison)

with host (numpy) buffers, i.e. exactly what the ISO_C_BINDING module passes, plus the device-resident ``dev_*`` calls
the native multi-GPU engine (:mod:`.dd`) drives.  Errors surface as :class:`lib.DlpError` carrying DL_POLY's error number.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from .lib import COREPART, DlpError, ptr

HALF_MINUS = np.nextafter(0.5, 0.0)          # constants.F90:198


class ShortRange:
    def __init__(self, device=0, dd=(1, 1, 1, 0, 0, 0)):
        self.L = _lib.load()
        h = C.c_void_p()
        rc = self.L.dlpgpu_create(C.byref(h), int(device))
        if rc != 0:
            raise DlpError(rc, "dlpgpu_create failed on device %d (is a CUDA GPU visible?)" % device)
        self.h = h
        self.device = device
        self.set_domain(dd)
        self.max_list = 0
        self.force_mode = 1

    def close(self):
        if getattr(self, "h", None):
            self.L.dlpgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise DlpError(rc, self.L.dlpgpu_last_error(self.h).decode(errors="replace"))

    # ---- setup -------------------------------------------------------------------------------------------------
    def set_domain(self, dd):
        a = np.ascontiguousarray(dd, dtype=np.int32)
        self.dd = tuple(int(v) for v in a)
        self._ck(self.L.dlpgpu_set_domain(self.h, ptr(a)))

    def set_cell(self, cell, imcon=1):
        c = np.ascontiguousarray(cell, dtype=np.float64).reshape(9)
        self.cell = c.copy()
        self.imcon = imcon
        self._ck(self.L.dlpgpu_set_cell(self.h, ptr(c), int(imcon)))

    def set_cutoffs(self, rcut, padding, pdplnc=50.0):
        self.rcut, self.padding = float(rcut), float(padding)
        self._ck(self.L.dlpgpu_set_cutoffs(self.h, float(rcut), float(padding), float(pdplnc)))

    def set_forcefield(self, ff):
        """ff: tables.ForceField (or anything with the same arrays -- what read_field leaves in vdw_type/electrostatic)."""
        lst = np.ascontiguousarray(ff.vdw_list_c, dtype=np.int32)
        ltp = np.ascontiguousarray(ff.ltp, dtype=np.int32)
        tp = np.ascontiguousarray(ff.tab_potential, dtype=np.float64)
        tf = np.ascontiguousarray(ff.tab_force, dtype=np.float64)
        par = np.ascontiguousarray(ff.param, dtype=np.float64)
        afs = np.ascontiguousarray(ff.afs, dtype=np.float64)
        bfs = np.ascontiguousarray(ff.bfs, dtype=np.float64)
        self._ck(self.L.dlpgpu_set_vdw(self.h, ff.ntypes, ptr(lst), ff.max_vdw, ff.n_vdw, ptr(ltp), ff.mxgrid, ptr(tp), ptr(tf),
                                       float(ff.rvdw), int(ff.force_shift), int(ff.direct), ptr(par), ptr(afs), ptr(bfs)))
        if ff.ew_active:
            e = np.ascontiguousarray(ff.erfc, dtype=np.float64)
            d = np.ascontiguousarray(ff.erfc_deriv, dtype=np.float64)
            self._ck(self.L.dlpgpu_set_ewald(self.h, 1, float(ff.alpha), float(ff.scaling), int(ff.ew_n), ptr(e), ptr(d),
                                             float(ff.ew_recip)))
        elif getattr(ff, "coul_kind", 0):
            rf = np.ascontiguousarray(ff.coul_rf, dtype=np.float64)
            if ff.coul_damp:
                e = np.ascontiguousarray(ff.erfc, dtype=np.float64)
                d = np.ascontiguousarray(ff.erfc_deriv, dtype=np.float64)
                self._ck(self.L.dlpgpu_set_coulomb(self.h, int(ff.coul_kind), 1, float(ff.scaling), float(ff.coul_force_shift),
                                                   float(ff.coul_energy_shift), ptr(rf), int(ff.ew_n), ptr(e), ptr(d), float(ff.ew_recip)))
            else:
                self._ck(self.L.dlpgpu_set_coulomb(self.h, int(ff.coul_kind), 0, float(ff.scaling), float(ff.coul_force_shift),
                                                   float(ff.coul_energy_shift), ptr(rf), 0, None, None, 0.0))
        else:
            self._ck(self.L.dlpgpu_set_ewald(self.h, 0, 0.0, 0.0, 0, None, None, 0.0))

    def set_force_mode(self, mode):
        self._ck(self.L.dlpgpu_set_force_mode(self.h, int(mode)))
        self.force_mode = int(mode)

    def set_pair_kernel(self, general_only=False, which=None):
        """Diagnostic: 0 automatic, 1 always the general pair kernel (the reference's operation order statement by statement)."""
        self._ck(self.L.dlpgpu_set_pair_kernel(self.h, int(which) if which is not None else int(bool(general_only))))

    def dev_xchg_last_ms(self):
        v = C.c_double(0.0)
        self._ck(self.L.dlpgpu_dev_xchg_last_ms(self.h, C.byref(v)))
        return v.value

    def dev_xchg_set_migration(self, scan_all):
        """Diagnostic: 1 = the migration stages of xchg_rebuild scan all atoms (see include/dlpgpu.h)."""
        self._ck(self.L.dlpgpu_dev_xchg_set_migration(self.h, int(scan_all)))

    def set_list_kernel(self, which):
        """Diagnostic: 0 trimmed candidate runs (default), 1 untrimmed, 2 trimmed + per-candidate ring (see include/dlpgpu.h)."""
        self._ck(self.L.dlpgpu_set_list_kernel(self.h, int(which)))

    def set_spme(self, kdim, nsplines=8):
        """SPME grid (k_vec_dim after adjust_kmax) and B-spline order; see include/dlpgpu.h."""
        k = np.ascontiguousarray(kdim, dtype=np.int32)
        self._ck(self.L.dlpgpu_set_spme(self.h, ptr(k), int(nsplines)))

    def dev_spme_forces(self, megatm):
        """ewald_spme_forces_coul on the device-resident atoms: adds the reciprocal forces, returns out[16]."""
        out = np.zeros(16)
        self._ck(self.L.dlpgpu_dev_spme_forces(self.h, int(megatm), ptr(out)))
        return out

    def dev_spme_spread(self, grid_ptr):
        """Stage 1 of the several-domain SPME: this rank's charges onto the caller's device grid (K1 K2 K3 doubles, zeroed here)."""
        self._ck(self.L.dlpgpu_dev_spme_spread(self.h, C.c_void_p(int(grid_ptr))))

    def dev_spme_solve_gather(self, grid_ptr):
        """Stage 2: the summed charge grid -> potential grid, forces of this rank's atoms; returns the rank's raw net force."""
        ft = np.zeros(3)
        self._ck(self.L.dlpgpu_dev_spme_solve_gather(self.h, C.c_void_p(int(grid_ptr)), ptr(ft)))
        return ft

    def dev_spme_finish(self, megatm, ftot_global, nranks):
        """Stage 3: net force of all ranks removed, forces added to the device arrays; this rank's share of out[16]."""
        out = np.zeros(16)
        ft = np.ascontiguousarray(ftot_global, dtype=np.float64)
        self._ck(self.L.dlpgpu_dev_spme_finish(self.h, int(megatm), ptr(ft), int(nranks), ptr(out)))
        return out

    def spme_forces(self, natms, parts, megatm):
        """Drop-in form: parts (host corePart array) up, reciprocal forces added to parts%f, records back; returns out[16]."""
        assert parts.dtype == COREPART and parts.flags.c_contiguous
        out = np.zeros(16)
        self._ck(self.L.dlpgpu_spme_forces(self.h, int(natms), ptr(parts), int(megatm), ptr(out)))
        return out

    def set_collect_pp(self, on=True):
        """stats%collect_pp: force calls also book per-particle energy / stress (see include/dlpgpu.h)."""
        self._ck(self.L.dlpgpu_set_collect_pp(self.h, int(on)))

    def get_pp(self, natms, pp_energy=None, pp_stress=None):
        """ADDS the per-particle sums of the last force call into (pp_energy(natms), pp_stress(natms, 9)); returns them."""
        e = np.zeros(natms) if pp_energy is None else pp_energy
        st = np.zeros((natms, 9)) if pp_stress is None else pp_stress
        self._ck(self.L.dlpgpu_get_pp(self.h, int(natms), ptr(e), ptr(st)))
        return e, st

    def pair_kernel_used(self):
        """(kernel of the last two_body_forces call: 1 general, 2 k_pair_v2; reserved)."""
        w, e = C.c_int(0), C.c_double(0.0)
        self._ck(self.L.dlpgpu_pair_kernel_used(self.h, C.byref(w), C.byref(e)))
        return w.value, e.value

    # ---- drop-in entry points (host buffers, Fortran index conventions) ------------------------------------------
    def link_cell_pairs(self, natms, nlast, parts, ltype, ltg, lfrzn=None, lbook=False, megfrz=0, list_excl=None,
                        max_list=None, want_list=True):
        """Returns the reference-format list as an int32 array (natms, max_list+4): column k+3 == list(k, i)."""
        assert parts.dtype == COREPART and parts.flags.c_contiguous
        ltype = np.ascontiguousarray(ltype, dtype=np.int32)
        ltg = np.ascontiguousarray(ltg, dtype=np.int32)
        lf = None if lfrzn is None else np.ascontiguousarray(lfrzn, dtype=np.int32)
        mxex = 0
        le = None
        if lbook and list_excl is not None:
            le = np.ascontiguousarray(list_excl, dtype=np.int32)
            mxex = le.shape[1] - 1
        self.max_list = int(max_list)
        out = np.zeros((natms, self.max_list + 4), dtype=np.int32) if want_list else None
        ibig = C.c_int(0)
        rc = self.L.dlpgpu_link_cell_pairs(self.h, int(natms), int(nlast), ptr(parts), ptr(ltype), ptr(ltg), ptr(lf), int(bool(lbook)),
                                           int(megfrz), int(mxex), ptr(le), self.max_list, ptr(out), C.byref(ibig))
        self.ibig = ibig.value
        self._ck(rc)
        return out

    def two_body_forces(self, natms, nlast, parts, unchanged_since_list=False):
        """Adds the pair forces into parts['fxx','fyy','fzz'][:natms]; returns the 16 partial sums of the C ABI.
        unchanged_since_list: the caller asserts parts has not been written since link_cell_pairs (skips the upload; with
        set_host_threads(n >= 1) only positions and charges matter, see include/dlpgpu.h)."""
        assert parts.dtype == COREPART and parts.flags.c_contiguous
        if unchanged_since_list:
            self._ck(self.L.dlpgpu_parts_unchanged_since_list(self.h))
        out = np.zeros(16)
        self._ck(self.L.dlpgpu_two_body_forces(self.h, int(natms), int(nlast), ptr(parts), ptr(out)))
        return out

    def set_host_threads(self, n):
        """0: whole corePart records by DMA (default); n >= 1: n host threads (de)interleave the useful fields (csrc/hostio.cu)."""
        self._ck(self.L.dlpgpu_set_host_threads(self.h, int(n)))

    def transfer_bytes(self, reset=False):
        """(h2d, d2h) bytes the drop-in entry points have copied over PCIe since the last reset."""
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        self._ck(self.L.dlpgpu_transfer_bytes(self.h, C.byref(a), C.byref(b), int(bool(reset))))
        return a.value, b.value

    def transfer_times(self):
        """Packed mode: host seconds in uploads, waiting for the force kernels, in downloads; and the call counts."""
        out = np.zeros(5)
        self._ck(self.L.dlpgpu_transfer_times(self.h, ptr(out)))
        return {"upload_s": out[0], "wait_s": out[1], "download_s": out[2], "uploads": int(out[3]), "downloads": int(out[4])}

    def rdf_collect(self, ntypes, rdf_list, n_pairs, max_grid, rdf=None):
        """rdf_collect + rdf_excl_collect on the device list; rdf: (n_pairs, max_grid) float64 counts, incremented."""
        lst = np.ascontiguousarray(rdf_list, dtype=np.int32)
        if rdf is None:
            rdf = np.zeros((n_pairs, max_grid))
        assert rdf.dtype == np.float64 and rdf.flags.c_contiguous and rdf.shape == (n_pairs, max_grid)
        self._ck(self.L.dlpgpu_rdf_collect(self.h, int(ntypes), ptr(lst), int(n_pairs), int(max_grid), ptr(rdf)))
        return rdf

    def vnl_check(self, natms, parts):
        tol = C.c_double(0.0)
        self._ck(self.L.dlpgpu_vnl_check(self.h, int(natms), ptr(parts), C.byref(tol)))
        return tol.value

    def vnl_set_check(self, nlast, parts):
        self._ck(self.L.dlpgpu_vnl_set_check(self.h, int(nlast), ptr(parts)))

    def vnl_update(self, tol_global):
        """neighbours.F90:182."""
        return tol_global >= HALF_MINUS * self.padding

    # ---- native device-resident mode -------------------------------------------------------------------------------
    def dev_setup_system(self, sysm, capacity=None):
        """Everything read_field / set_bounds would hand over for ``sysm`` (a systems.System)."""
        self.set_cell(sysm.cell, sysm.imcon)
        self.set_cutoffs(sysm.rcut, sysm.padding, sysm.pdplnc)
        self.set_forcefield(sysm.ff)
        t = np.ascontiguousarray(sysm.type_site, dtype=np.int32)
        q = np.ascontiguousarray(sysm.charge_site, dtype=np.float64)
        f = np.ascontiguousarray(sysm.freeze_site, dtype=np.int32)
        w = np.ascontiguousarray(sysm.weight_site, dtype=np.float64)
        self._ck(self.L.dlpgpu_dev_set_sites(self.h, len(t), ptr(t), ptr(q), ptr(f), ptr(w)))
        if sysm.excl is not None:
            e = np.ascontiguousarray(sysm.excl, dtype=np.int32)
            self._ck(self.L.dlpgpu_dev_set_excl(self.h, e.shape[0], e.shape[1] - 1, ptr(e)))
        self.max_list = int(sysm.max_list)
        self._ck(self.L.dlpgpu_dev_set_list_capacity(self.h, self.max_list, int(sysm.megfrz)))

    def dev_load_atoms(self, xyz, vel, ltg, lsite, capacity=0):
        x = np.ascontiguousarray(xyz, dtype=np.float64)
        v = None if vel is None else np.ascontiguousarray(vel, dtype=np.float64)
        g = np.ascontiguousarray(ltg, dtype=np.int32)
        s = np.ascontiguousarray(lsite, dtype=np.int32)
        self._ck(self.L.dlpgpu_dev_load_atoms(self.h, x.shape[0], ptr(x), ptr(v), ptr(g), ptr(s), int(capacity)))

    def dev_set_halo_width(self, ecw):
        e = np.ascontiguousarray(ecw, dtype=np.float64)
        self._ck(self.L.dlpgpu_dev_set_halo_width(self.h, ptr(e)))

    def dev_counts(self):
        a, b = C.c_int(), C.c_int()
        self._ck(self.L.dlpgpu_dev_counts(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def dev_zero_forces(self):
        self._ck(self.L.dlpgpu_dev_zero_forces(self.h))

    def dev_vv(self, stage, dt):
        self._ck(self.L.dlpgpu_dev_vv(self.h, int(stage), float(dt)))

    def dev_vnl_check(self):
        tol = C.c_double(0.0)
        self._ck(self.L.dlpgpu_dev_vnl_check(self.h, C.byref(tol)))
        return tol.value

    def dev_halo_serial(self):
        self._ck(self.L.dlpgpu_dev_halo_serial(self.h))

    def dev_refresh_serial(self):
        self._ck(self.L.dlpgpu_dev_refresh_serial(self.h))

    def dev_relocate_serial(self):
        self._ck(self.L.dlpgpu_dev_relocate_serial(self.h))

    def dev_halo_begin(self):
        self._ck(self.L.dlpgpu_dev_halo_begin(self.h))

    def dev_halo_pack(self, mdir, sendbuf_ptr, capacity_atoms):
        n = C.c_int(0)
        rc = self.L.dlpgpu_dev_halo_pack(self.h, int(mdir), C.c_void_p(sendbuf_ptr), int(capacity_atoms), C.byref(n))
        if rc != 0 and rc != 54:
            self._ck(rc)
        return rc, n.value

    def dev_halo_unpack(self, mdir, recvbuf_ptr, count):
        self._ck(self.L.dlpgpu_dev_halo_unpack(self.h, int(mdir), C.c_void_p(recvbuf_ptr), int(count)))

    def dev_halo_end(self):
        self._ck(self.L.dlpgpu_dev_halo_end(self.h))

    def dev_p2p_init(self, rank, nranks, capacity_atoms):
        """Returns the handle blob (DLPGPU_P2P_BLOB bytes) of this rank's two peer-visible coordinate buffers."""
        h = np.zeros(192, dtype=np.uint8)
        self._ck(self.L.dlpgpu_dev_p2p_init(self.h, int(rank), int(nranks), int(capacity_atoms), ptr(h)))
        return h

    def dev_p2p_open(self, all_handles):
        a = np.ascontiguousarray(all_handles, dtype=np.uint8)
        self._ck(self.L.dlpgpu_dev_p2p_open(self.h, ptr(a)))

    def dev_publish(self):
        self._ck(self.L.dlpgpu_dev_publish(self.h))

    def dev_refresh_pull(self):
        self._ck(self.L.dlpgpu_dev_refresh_pull(self.h))

    def dev_xchg_init(self, rank, nranks, cap_reloc_atoms, cap_halo_atoms):
        """Allocates the peer-visible exchange region (gmax mailboxes + per-stage receive buffers); returns its handle blob
        (DLPGPU_XCHG_BLOB bytes)."""
        h = np.zeros(128, dtype=np.uint8)
        self._ck(self.L.dlpgpu_dev_xchg_init(self.h, int(rank), int(nranks), int(cap_reloc_atoms), int(cap_halo_atoms), ptr(h)))
        return h

    def dev_xchg_open(self, all_handles):
        a = np.ascontiguousarray(all_handles, dtype=np.uint8)
        self._ck(self.L.dlpgpu_dev_xchg_open(self.h, ptr(a)))

    def dev_xchg_rebuild(self, neigh, seq):
        """relocate_particles + set_halo_particles + vnl_set_check as one device-side exchange; returns (natms, nlast)."""
        nb = np.ascontiguousarray(neigh, dtype=np.int32)
        a, b = C.c_int(0), C.c_int(0)
        self._ck(self.L.dlpgpu_dev_xchg_rebuild(self.h, ptr(nb), C.c_ulonglong(int(seq)), C.byref(a), C.byref(b)))
        return a.value, b.value

    def dev_xchg_gmax(self, seq):
        """vnl_check + gmax: the largest displacement over all ranks, reduced by the GPUs through their mailboxes."""
        tol = C.c_double(0.0)
        self._ck(self.L.dlpgpu_dev_xchg_gmax(self.h, C.c_ulonglong(int(seq)), C.byref(tol)))
        return tol.value

    def dev_set_rebuild_every(self, every):
        self._ck(self.L.dlpgpu_dev_set_rebuild_every(self.h, int(every)))

    def dev_xchg_set_timeout(self, seconds):
        self._ck(self.L.dlpgpu_dev_xchg_set_timeout(self.h, float(seconds)))

    def dev_md_step(self, neigh, dt, gseq, rseq):
        """One MD step enqueued from C (see dlpgpu_dev_md_step); returns (rebuilt, previous step's sums or None, list_ms)."""
        nb = self._neigh_cache.get(tuple(neigh)) if hasattr(self, "_neigh_cache") else None
        if nb is None:
            if not hasattr(self, "_neigh_cache"):
                self._neigh_cache = {}
                self._step_out = np.zeros(16)
            nb = np.ascontiguousarray(neigh, dtype=np.int32)
            self._neigh_cache[tuple(neigh)] = nb
        reb, have, lms = C.c_int(0), C.c_int(0), C.c_double(0.0)
        self._ck(self.L.dlpgpu_dev_md_step(self.h, ptr(nb), float(dt), C.c_ulonglong(int(gseq)), C.c_ulonglong(int(rseq)), C.byref(reb),
                                           ptr(self._step_out), C.byref(have), C.byref(lms)))
        return bool(reb.value), (self._step_out.copy() if have.value else None), lms.value

    def dev_halo_stage_counts(self):
        a, b = np.zeros(6, dtype=np.int32), np.zeros(6, dtype=np.int32)
        self._ck(self.L.dlpgpu_dev_halo_stage_counts(self.h, ptr(a), ptr(b)))
        return [int(v) for v in a], [int(v) for v in b]

    def dev_refresh_pack(self, mdir, sendbuf_ptr):
        n = C.c_int(0)
        self._ck(self.L.dlpgpu_dev_refresh_pack(self.h, int(mdir), C.c_void_p(sendbuf_ptr), C.byref(n)))
        return n.value

    def dev_refresh_unpack(self, mdir, recvbuf_ptr, count):
        self._ck(self.L.dlpgpu_dev_refresh_unpack(self.h, int(mdir), C.c_void_p(recvbuf_ptr), int(count)))

    def dev_relocate_begin(self):
        self._ck(self.L.dlpgpu_dev_relocate_begin(self.h))

    def dev_relocate_pack(self, mdir, sendbuf_ptr, capacity_atoms):
        n = C.c_int(0)
        rc = self.L.dlpgpu_dev_relocate_pack(self.h, int(mdir), C.c_void_p(sendbuf_ptr), int(capacity_atoms), C.byref(n))
        if rc != 0 and rc != 54:
            self._ck(rc)
        return rc, n.value

    def dev_relocate_unpack(self, mdir, recvbuf_ptr, count):
        self._ck(self.L.dlpgpu_dev_relocate_unpack(self.h, int(mdir), C.c_void_p(recvbuf_ptr), int(count)))

    def dev_relocate_end(self):
        n = C.c_int(0)
        self._ck(self.L.dlpgpu_dev_relocate_end(self.h, C.byref(n)))
        return n.value

    def dev_link_cell_pairs(self, want_ref_list=False):
        ibig = C.c_int(0)
        rc = self.L.dlpgpu_dev_link_cell_pairs(self.h, int(bool(want_ref_list)), C.byref(ibig))
        self.ibig = ibig.value
        self._ck(rc)

    def dev_two_body_forces(self, zero_forces=True):
        out = np.zeros(16)
        self._ck(self.L.dlpgpu_dev_two_body_forces(self.h, int(bool(zero_forces)), ptr(out)))
        return out

    def dev_two_body_forces_async(self, zero_forces=True):
        """Enqueues two_body_forces without waiting for the sums; collect them with dev_fetch_results()."""
        self._ck(self.L.dlpgpu_dev_two_body_forces(self.h, int(bool(zero_forces)), None))

    def dev_fetch_results(self):
        out = np.zeros(16)
        self._ck(self.L.dlpgpu_dev_fetch_results(self.h, ptr(out)))
        return out

    def dev_list_pairs(self):
        n = C.c_longlong(0)
        self._ck(self.L.dlpgpu_dev_list_pairs(self.h, C.byref(n)))
        return int(n.value)

    # ---- read-back -------------------------------------------------------------------------------------------------
    def dev_get_parts(self, n=None):
        if n is None:
            n = self.dev_counts()[1]
        p = np.zeros(n, dtype=COREPART)
        self._ck(self.L.dlpgpu_dev_get_parts(self.h, ptr(p), int(n)))
        return p

    def dev_get_ints(self, n=None):
        if n is None:
            n = self.dev_counts()[1]
        arrs = [np.zeros(n, dtype=np.int32) for _ in range(5)]
        self._ck(self.L.dlpgpu_dev_get_ints(self.h, int(n), *[ptr(a) for a in arrs]))
        return dict(zip(["ltg", "lsite", "ltype", "lfrzn", "ixyz"], arrs))

    def dev_get_vel(self, n=None):
        if n is None:
            n = self.dev_counts()[0]
        v = np.zeros((n, 3))
        self._ck(self.L.dlpgpu_dev_get_vel(self.h, int(n), ptr(v)))
        return v

    def dev_get_list(self):
        natms = self.dev_counts()[0]
        out = np.zeros((natms, self.max_list + 4), dtype=np.int32)
        self._ck(self.L.dlpgpu_dev_get_list(self.h, natms, self.max_list, ptr(out)))
        return out

    def dev_get_cells(self, arrays=True):
        info = np.zeros(6, dtype=np.int32)
        self._ck(self.L.dlpgpu_dev_get_cells(self.h, ptr(info), None, None, None))
        keys = ["nlx", "nly", "nlz", "nlp", "ncells", "nsbcll"]
        d = dict(zip(keys, (int(v) for v in info)))
        if arrays:
            nlast = self.dev_counts()[1]
            wc = np.zeros(nlast, dtype=np.int32)
            al = np.zeros(nlast, dtype=np.int32)
            ls = np.zeros(d["ncells"] + 2, dtype=np.int32)
            self._ck(self.L.dlpgpu_dev_get_cells(self.h, ptr(info), ptr(wc), ptr(al), ptr(ls)))
            d.update(which_cell=wc, at_list=al, lct_start=ls)
        return d

    def dev_get_full_row(self, i):
        cap = self.max_list + 64
        m = np.zeros(cap, dtype=np.int32)
        x = np.zeros(cap, dtype=np.int32)
        nm, nx = C.c_int(), C.c_int()
        self._ck(self.L.dlpgpu_dev_get_full_row(self.h, int(i), C.byref(nm), ptr(m), C.byref(nx), ptr(x), cap))
        return m[:nm.value].copy(), x[:nx.value].copy()

    # ---- measurement -----------------------------------------------------------------------------------------------
    def fp64_peak(self, seconds=0.2):
        t = C.c_double(0.0)
        self._ck(self.L.dlpgpu_fp64_peak(self.h, float(seconds), C.byref(t)))
        return t.value

    def last_timings(self):
        t = np.zeros(4)
        self._ck(self.L.dlpgpu_last_timings(self.h, ptr(t)))
        return dict(list_ms=t[0], force_ms=t[1], pair_kernel_ms=t[2], full_list_kernel_ms=t[3])

    def launch_count(self):
        return int(self.L.dlpgpu_launch_count(self.h))

    def stream(self):
        return self.L.dlpgpu_stream(self.h)
