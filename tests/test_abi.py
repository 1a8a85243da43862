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


def test_fortran_binding_agrees_with_the_header():
    """fortran/dlp_gpu_binding.F90 cannot be compiled in this image, so hold its Bind(C) interfaces against include/dlpgpu.h
    textually: every bound name is declared in the header with the same number of arguments, value / reference passing
    matches (scalars by Value, arrays and out-arguments by reference), and the drop-in entry points are all bound."""
    hdr = open(os.path.join(ROOT, "include", "dlpgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|const char\*|long long|void\*)\s+(dlpgpu_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S):
        args = [a.strip() for a in m.group(2).replace("\n", " ").split(",")]
        protos[m.group(1)] = [] if args == ["void"] else args
    src = open(os.path.join(ROOT, "fortran", "dlp_gpu_binding.F90")).read()
    joined = re.sub(r"&\s*\n\s*", " ", src)                                   # continuation lines
    bound = {}
    for m in re.finditer(r"Function\s+(\w+)\s*\(([^)]*)\)\s*Bind\(C,\s*name='(\w+)'\)(.*?)End Function", joined, flags=re.S):
        fname, fargs, cname, body = m.group(1), [a.strip() for a in m.group(2).split(",") if a.strip()], m.group(3), m.group(4)
        assert fname == cname
        bound[cname] = (fargs, body)
    for must in ("dlpgpu_create", "dlpgpu_destroy", "dlpgpu_set_domain", "dlpgpu_set_cell", "dlpgpu_set_cutoffs", "dlpgpu_set_vdw",
                 "dlpgpu_set_ewald", "dlpgpu_set_coulomb", "dlpgpu_link_cell_pairs", "dlpgpu_two_body_forces", "dlpgpu_vnl_check",
                 "dlpgpu_rdf_collect", "dlpgpu_parts_unchanged_since_list"):
        assert must in bound, must
    for cname, (fargs, body) in bound.items():
        assert cname in protos, "%s is not declared in dlpgpu.h" % cname
        cargs = protos[cname]
        assert len(fargs) == len(cargs), (cname, fargs, cargs)
        for fa, ca in zip(fargs, cargs):
            decl = [ln for ln in body.split("\n") if re.search(r"::.*\b%s\b" % re.escape(fa), ln)]
            assert decl, (cname, fa)
            by_value = "Value" in decl[0]
            c_is_pointer = "*" in ca or "[" in ca
            if "dlpgpu_ctx**" in ca.replace(" ", ""):
                assert not by_value                                            # Type(c_ptr), Intent(Out)
            elif "dlpgpu_ctx*" in ca.replace(" ", ""):
                assert by_value and "c_ptr" in decl[0], (cname, fa, decl[0])   # the handle itself
            elif c_is_pointer:
                # arrays / out-arguments by reference, or a C pointer passed as Type(c_ptr), Value (c_loc of a Fortran array)
                assert (not by_value) or "c_ptr" in decl[0], (cname, fa, decl[0])
            else:
                assert by_value, (cname, fa, decl[0])                          # C scalars
                want = "c_double" if ca.split()[0] == "double" else "c_int"
                assert want in decl[0], (cname, fa, decl[0])
