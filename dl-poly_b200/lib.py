"""ctypes binding of libdlpgpu.so -- every symbol include/dlpgpu.h declares.

There is no CPU fallback: if the shared library is missing or cannot be loaded this module raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdlpgpu.so")

COREPART = np.dtype([("xxx", "f8"), ("yyy", "f8"), ("zzz", "f8"), ("fxx", "f8"), ("fyy", "f8"), ("fzz", "f8"),
                     ("chge", "f8"), ("pad1", "i4"), ("pad2", "i4")])       # particle.F90:14-20
assert COREPART.itemsize == 64

vp, ci, cd = C.c_void_p, C.c_int, C.c_double
pi_, pd_ = C.POINTER(C.c_int), C.POINTER(C.c_double)

# name -> (restype, argtypes); the keys are exactly the functions declared in include/dlpgpu.h
SIGNATURES = {
    "dlpgpu_create": (ci, [C.POINTER(vp), ci]),
    "dlpgpu_destroy": (ci, [vp]),
    "dlpgpu_last_error": (C.c_char_p, [vp]),
    "dlpgpu_version": (ci, []),
    "dlpgpu_launch_count": (C.c_longlong, [vp]),
    "dlpgpu_stream": (vp, [vp]),
    "dlpgpu_set_domain": (ci, [vp, vp]),
    "dlpgpu_set_cell": (ci, [vp, vp, ci]),
    "dlpgpu_set_cutoffs": (ci, [vp, cd, cd, cd]),
    "dlpgpu_set_vdw": (ci, [vp, ci, vp, ci, ci, vp, ci, vp, vp, cd, ci, ci, vp, vp, vp]),
    "dlpgpu_set_ewald": (ci, [vp, ci, cd, cd, ci, vp, vp, cd]),
    "dlpgpu_set_coulomb": (ci, [vp, ci, ci, cd, cd, cd, vp, ci, vp, vp, cd]),
    "dlpgpu_link_cell_pairs": (ci, [vp, ci, ci, vp, vp, vp, vp, ci, ci, ci, vp, ci, vp, pi_]),
    "dlpgpu_two_body_forces": (ci, [vp, ci, ci, vp, vp]),
    "dlpgpu_parts_unchanged_since_list": (ci, [vp]),
    "dlpgpu_set_host_threads": (ci, [vp, ci]),
    "dlpgpu_transfer_bytes": (ci, [vp, vp, vp, ci]),
    "dlpgpu_transfer_times": (ci, [vp, vp]),
    "dlpgpu_rdf_collect": (ci, [vp, ci, vp, ci, ci, vp]),
    "dlpgpu_vnl_check": (ci, [vp, ci, vp, pd_]),
    "dlpgpu_vnl_set_check": (ci, [vp, ci, vp]),
    "dlpgpu_dev_set_sites": (ci, [vp, ci, vp, vp, vp, vp]),
    "dlpgpu_dev_set_excl": (ci, [vp, ci, ci, vp]),
    "dlpgpu_dev_set_halo_width": (ci, [vp, vp]),
    "dlpgpu_dev_set_list_capacity": (ci, [vp, ci, ci]),
    "dlpgpu_dev_load_atoms": (ci, [vp, ci, vp, vp, vp, vp, ci]),
    "dlpgpu_dev_counts": (ci, [vp, pi_, pi_]),
    "dlpgpu_dev_zero_forces": (ci, [vp]),
    "dlpgpu_dev_vv": (ci, [vp, ci, cd]),
    "dlpgpu_dev_vnl_check": (ci, [vp, pd_]),
    "dlpgpu_dev_halo_begin": (ci, [vp]),
    "dlpgpu_dev_halo_pack": (ci, [vp, ci, vp, ci, pi_]),
    "dlpgpu_dev_halo_unpack": (ci, [vp, ci, vp, ci]),
    "dlpgpu_dev_halo_end": (ci, [vp]),
    "dlpgpu_dev_refresh_pack": (ci, [vp, ci, vp, pi_]),
    "dlpgpu_dev_refresh_unpack": (ci, [vp, ci, vp, ci]),
    "dlpgpu_dev_p2p_init": (ci, [vp, ci, ci, ci, vp]),
    "dlpgpu_dev_p2p_open": (ci, [vp, vp]),
    "dlpgpu_dev_publish": (ci, [vp]),
    "dlpgpu_dev_refresh_pull": (ci, [vp]),
    "dlpgpu_dev_xchg_init": (ci, [vp, ci, ci, ci, ci, vp]),
    "dlpgpu_dev_xchg_open": (ci, [vp, vp]),
    "dlpgpu_dev_xchg_set_timeout": (ci, [vp, cd]),
    "dlpgpu_dev_set_rebuild_every": (ci, [vp, ci]),
    "dlpgpu_dev_xchg_rebuild": (ci, [vp, vp, C.c_ulonglong, pi_, pi_]),
    "dlpgpu_dev_xchg_gmax": (ci, [vp, C.c_ulonglong, pd_]),
    "dlpgpu_dev_md_step": (ci, [vp, vp, cd, C.c_ulonglong, C.c_ulonglong, pi_, vp, pi_, pd_]),
    "dlpgpu_dev_halo_stage_counts": (ci, [vp, vp, vp]),
    "dlpgpu_dev_halo_serial": (ci, [vp]),
    "dlpgpu_dev_refresh_serial": (ci, [vp]),
    "dlpgpu_dev_relocate_serial": (ci, [vp]),
    "dlpgpu_dev_relocate_begin": (ci, [vp]),
    "dlpgpu_dev_relocate_pack": (ci, [vp, ci, vp, ci, pi_]),
    "dlpgpu_dev_relocate_unpack": (ci, [vp, ci, vp, ci]),
    "dlpgpu_dev_relocate_end": (ci, [vp, pi_]),
    "dlpgpu_dev_link_cell_pairs": (ci, [vp, ci, pi_]),
    "dlpgpu_dev_two_body_forces": (ci, [vp, ci, vp]),
    "dlpgpu_dev_fetch_results": (ci, [vp, vp]),
    "dlpgpu_dev_list_pairs": (ci, [vp, C.POINTER(C.c_longlong)]),
    "dlpgpu_dev_get_parts": (ci, [vp, vp, ci]),
    "dlpgpu_dev_get_ints": (ci, [vp, ci, vp, vp, vp, vp, vp]),
    "dlpgpu_dev_get_vel": (ci, [vp, ci, vp]),
    "dlpgpu_dev_get_list": (ci, [vp, ci, ci, vp]),
    "dlpgpu_dev_get_cells": (ci, [vp, vp, vp, vp, vp]),
    "dlpgpu_dev_get_full_row": (ci, [vp, ci, pi_, vp, pi_, vp, ci]),
    "dlpgpu_fp64_peak": (ci, [vp, cd, pd_]),
    "dlpgpu_last_timings": (ci, [vp, vp]),
    "dlpgpu_set_force_mode": (ci, [vp, ci]),
    "dlpgpu_set_pair_kernel": (ci, [vp, ci]),
    "dlpgpu_set_list_kernel": (ci, [vp, ci]),
    "dlpgpu_dev_xchg_set_migration": (ci, [vp, ci]),
    "dlpgpu_dev_xchg_last_ms": (ci, [vp, vp]),
    "dlpgpu_set_spme": (ci, [vp, vp, ci]),
    "dlpgpu_dev_spme_forces": (ci, [vp, ci, vp]),
    "dlpgpu_spme_forces": (ci, [vp, ci, vp, ci, vp]),
    "dlpgpu_dev_spme_spread": (ci, [vp, vp]),
    "dlpgpu_dev_spme_solve_gather": (ci, [vp, vp, vp]),
    "dlpgpu_dev_spme_finish": (ci, [vp, ci, vp, ci, vp]),
    "dlpgpu_set_collect_pp": (ci, [vp, ci]),
    "dlpgpu_get_pp": (ci, [vp, ci, vp, vp]),
    "dlpgpu_pair_kernel_used": (ci, [vp, pi_, pd_]),
}

_lib = None


def load():
    """Loads libdlpgpu.so (raises if absent -- the product path has no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libdlpgpu.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback for the DL_POLY short-range path" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)       # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


class DlpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dlpgpu error %d: %s" % (code, msg))
        self.code = code


def ptr(a):
    if a is None:
        return None
    return C.c_void_p(a.ctypes.data)
