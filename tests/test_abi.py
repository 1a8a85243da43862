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
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not bad.search(txt), "%s references the oracle" % f


def test_cpp_host_fails_loudly_without_gpu(tmp_path):
    """The C++ host side has no CPU fallback either: without a GPU gpu_short_range's constructor raises and the check program
    exits non-zero with the message."""
    import subprocess
    import struct
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dl_poly_b200 import build
    exe = build.build_host()
    fin = tmp_path / "in.bundle"
    import numpy as np

    def rec(name, arr):
        a = np.ascontiguousarray(arr)
        dt = b"d" if a.dtype.kind == "f" else b"i"
        a = a.astype(np.float64 if dt == b"d" else np.int32)
        return name.encode().ljust(24, b"\0") + dt + b"\0" * 7 + struct.pack("<q", a.size) + a.tobytes()

    recs = [rec("cell", np.diag([30.0, 30.0, 30.0]).reshape(9)), rec("imcon", [1]), rec("megatm", [10]), rec("rcut", [8.0]),
            rec("rvdw", [8.0]), rec("padding", [0.2]), rec("ntypes", [1]), rec("force_shift", [0]), rec("direct", [0]),
            rec("pot_pairs", np.array([1, 1, 2])), rec("pot_param", [99.61, 3.405, 0, 0, 0, 0, 0.0]), rec("electro_key", [0]),
            rec("eps", [1.0]), rec("damping", [0.0]), rec("natms", [0]), rec("nlast", [0])]
    fin.write_bytes(b"".join(recs))
    r = subprocess.run([exe, "md", str(fin), str(tmp_path / "out.bundle")], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
