"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol include/dlpgpu.h declares.
No compute calls are made (there is no GPU in the CPU container)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "dlpgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dlpgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_functions()
    for must in ["dlpgpu_create", "dlpgpu_destroy", "dlpgpu_last_error", "dlpgpu_set_vdw", "dlpgpu_set_ewald",
                 "dlpgpu_link_cell_pairs", "dlpgpu_two_body_forces", "dlpgpu_vnl_check", "dlpgpu_dev_halo_pack",
                 "dlpgpu_dev_refresh_pack", "dlpgpu_dev_relocate_pack"]:
        assert must in names


def test_library_loads_and_exports_every_declared_symbol():
    from dl_poly_b200 import build, lib
    build.build()
    L = lib.load()
    names = declared_functions()
    assert sorted(lib.SIGNATURES) == names, "python binding and header disagree"
    for n in names:
        assert hasattr(L, n), "libdlpgpu.so does not export %s" % n
    assert L.dlpgpu_version() >= 100


def test_corepart_layout_is_64_bytes():
    from dl_poly_b200 import lib
    assert lib.COREPART.itemsize == 64
    assert [lib.COREPART.fields[k][1] for k in ("xxx", "yyy", "zzz", "fxx", "fyy", "fzz", "chge", "pad1", "pad2")] == \
        [0, 8, 16, 24, 32, 40, 48, 56, 60]


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: on a box without a GPU the context cannot be created and the host mirror raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dl_poly_b200 import engine, lib
    with pytest.raises(lib.DlpError):
        engine.ShortRange(device=0)


def test_product_never_imports_the_oracle():
    """The product package must not import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "dl-poly_b200")
    bad = re.compile(r"^\s*(from|import)\s+\S*oracle|libdlp_oracle|dlp_oracle\.|oracle/|ora_world|ora_dom", re.M)
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), "%s references the oracle" % f
