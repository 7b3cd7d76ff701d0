"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/maestro_b200.h declares, and refuses to compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from maestro_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "maestro_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mgpu_[a-zA-Z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    l = lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(l, s), "libmaestro_b200.so does not export %s" % s


def test_python_abi_mirror_matches_header():
    assert sorted(abi.all_symbols()) == header_symbols()


def test_struct_layout_matches_c(tmp_path):
    """sizeof / offsetof of every struct, as gcc lays out include/maestro_b200.h, against the ctypes mirror"""
    import subprocess

    src = tmp_path / "layout.c"
    fields = {"mgpu_fab": [f[0] for f in abi.mgpu_fab._fields_], "mgpu_params": [f[0] for f in abi.mgpu_params._fields_],
              "mgpu_halo_plan": [f[0] for f in abi.mgpu_halo_plan._fields_],
              "mgpu_geom": [f[0] for f in abi.mgpu_geom._fields_], "mgpu_eos": [f[0] for f in abi.mgpu_eos._fields_]}
    body = ['#include <stdio.h>', '#include <stddef.h>', '#include "maestro_b200.h"', 'int main(void) {']
    for st, fs in fields.items():
        body.append('printf("%s %%zu\\n", sizeof(%s));' % (st, st))
        for f in fs:
            body.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (st, f, st, f))
    body.append('return 0; }')
    src.write_text("\n".join(body))
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for st, fs in fields.items():
        cls = getattr(abi, st)
        assert C.sizeof(cls) == int(out[st]), st
        for f in fs:
            assert getattr(cls, f).offset == int(out["%s.%s" % (st, f)]), (st, f)


def test_fortran_shim_binds_only_declared_symbols():
    """every bind(C, name="...") of shim/maestro_b200_shim.f90 is an entry point the header declares (and the library
    exports), and every operator entry point of the header has a binding in the shim"""
    txt = open(os.path.join(ROOT, "shim", "maestro_b200_shim.f90")).read()
    bound = set(re.findall(r'bind\(C,\s*name="(mgpu_[A-Za-z0-9_]+)"\)', txt))
    declared = set(header_symbols())
    assert bound, "no bindings found"
    assert bound <= declared, sorted(bound - declared)
    # operators (not the test / profiling helpers) must all be reachable from Fortran
    helpers = {s for s in declared if any(k in s for k in ("profile", "launch_count", "copy_bytes", "version", "stream",
                                                           "halo_plan", "synchronize", "host_unregister"))}
    missing = declared - bound - helpers
    assert not missing, sorted(missing)


def test_fortran_shim_struct_mirrors_match():
    """the bind(C) derived types of the shim list the same fields, in the same order and with the same kinds and
    extents, as the C structs (through the ctypes mirror, itself checked against gcc's layout above)"""
    txt = open(os.path.join(ROOT, "shim", "maestro_b200_shim.f90")).read()
    kinds = {"integer(c_int)": C.c_int, "real(c_double)": C.c_double, "type(c_ptr)": None}
    for st in ("mgpu_fab", "mgpu_params", "mgpu_geom", "mgpu_eos"):
        body = re.search(r"type, bind\(C\), public :: %s\n(.*?)end type %s" % (st, st), txt, flags=re.S).group(1)
        got = []
        for line in body.splitlines():
            line = line.split("!")[0].strip()
            if not line:
                continue
            kind, names = [x.strip() for x in line.split("::")]
            for nm in re.findall(r"([A-Za-z_0-9]+)(?:\((\d+)\))?", names):
                got.append((nm[0], kind, int(nm[1]) if nm[1] else 1))
        want = getattr(abi, st)._fields_
        assert [g[0] for g in got] == [w[0] for w in want], st
        for (name, kind, ext), (wname, wtype) in zip(got, want):
            n = getattr(wtype, "_length_", 1)
            base = getattr(wtype, "_type_", wtype) if n > 1 else wtype
            assert ext == n, (st, name)
            if kinds[kind] is None:
                assert C.sizeof(base) == C.sizeof(C.c_void_p), (st, name)
            else:
                assert base is kinds[kind], (st, name)


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    l = lib.load()
    assert l.mgpu_init(0) != 0
    assert b"no CUDA device" in l.mgpu_last_error()
    from maestro_b200 import Fab, face_fabs, make_params, make_adv_bc

    p = make_params(2, n=[8, 8, 1])
    s = Fab([0, 0, 0], [7, 7, 0], 4, p.nscal, dm=2)
    f = Fab([0, 0, 0], [7, 7, 0], 1, p.nscal, dm=2)
    with pytest.raises(RuntimeError, match="not initialised"):
        lib.ops().make_edge_scal(p, s, face_fabs([0, 0, 0], [7, 7, 0], 0, p.nscal, 2),
                                 face_fabs([0, 0, 0], [7, 7, 0], 1, 1, 2), f,
                                 make_adv_bc(p, [[-1, -1], [-1, -1]]), False, 1, 3, 1, False)


def test_fortran_e_format():
    """e17.10 as gfortran prints it (the reference's formats 2000-2003)"""
    from maestro_b200.operators import fortran_e

    assert fortran_e(1.0) == " 0.1000000000E+01"
    assert fortran_e(-0.105604113268602) == "-0.1056041133E+00"
    assert fortran_e(0.0) == " 0.0000000000E+00"
    assert fortran_e(9.99999999999e9) == " 0.1000000000E+11"
    assert fortran_e(1.0e-10) == " 0.1000000000E-09"
    assert fortran_e(-2.5e120) == "-0.2500000000+121"


def _c_prototypes():
    """name -> (return type, [("val" | "ptr", base type, number of stars)]) from include/maestro_b200.h"""
    txt = open(os.path.join(ROOT, "include", "maestro_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9 \*]*?)\b(mgpu_[A-Za-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        params = []
        if args and args != "void":
            for a in args.split(","):
                stars = a.count("*")
                base = re.sub(r"\bconst\b|\*|\bunsigned\b|\bstruct\b", " ", a).split()
                params.append(("ptr" if stars else "val", base[0], stars))
        out[name] = (ret, params)
    return out


def _fortran_interfaces():
    """C name -> (result kind, [(dummy, (type, has `value`, is an array))]) from the interface blocks of the shim"""
    txt = open(os.path.join(ROOT, "shim", "maestro_b200_shim.f90")).read()
    txt = re.sub(r"&\s*\n\s*&?", " ", txt)  # continuation lines
    out = {}
    pat = (r"^\s*(?:(integer\(c_int\)|integer\(c_long\)|type\(c_ptr\)|real\(c_double\))\s+function|subroutine)\s+(\w+)"
           r"\s*\(([^)]*)\)\s*bind\(C,\s*name=\"(mgpu_\w+)\"\)(.*?)end (?:function|subroutine)")
    for m in re.finditer(pat, txt, flags=re.S | re.M | re.I):
        ret, _, dummies, cname, body = m.groups()
        dummies = [d.strip() for d in dummies.split(",") if d.strip()]
        decl = {}
        for line in body.splitlines():
            line = line.split("!")[0].strip()
            if "::" not in line or line.lower().startswith("import"):
                continue
            left, names = line.split("::")
            attrs = [x.strip().lower() for x in re.split(r",(?![^()]*\))", left)]
            for nm in re.findall(r"([A-Za-z_]\w*)(\([^)]*\))?", names):
                decl[nm[0].lower()] = (attrs[0], "value" in attrs, bool(nm[1]))
        out[cname] = (ret, [(d, decl.get(d.lower())) for d in dummies])
    return out


def test_fortran_shim_argument_lists_match_header():
    """every interface block of the shim against the C prototype it binds: the same number of arguments in the same
    order, by value exactly where C passes by value (with the matching kind), by reference where C takes a pointer
    (with the matching element type; `type(c_ptr), value` stands for any pointer and an array of `type(c_ptr)` for a
    pointer to pointers), and the matching result kind.  No Fortran compiler is available to do this check."""
    cp, fi = _c_prototypes(), _fortran_interfaces()
    assert len(fi) >= 70
    val_kind = {"int": "integer(c_int)", "long": "integer(c_long)", "double": "real(c_double)"}
    ref_kind = {"int": "integer(c_int)", "double": "real(c_double)", "long": "integer(c_long)", "char": "character(kind=c_char)",
                "mgpu_params": "type(mgpu_params)", "mgpu_fab": "type(mgpu_fab)", "mgpu_geom": "type(mgpu_geom)",
                "mgpu_eos": "type(mgpu_eos)"}
    ret_kind = {"int": "integer(c_int)", "long": "integer(c_long)", "const char*": "type(c_ptr)", "double": "real(c_double)"}
    for name, (fret, fargs) in sorted(fi.items()):
        cret, cargs = cp[name]
        assert fret is not None and fret.lower() == ret_kind[cret], (name, fret, cret)
        assert len(fargs) == len(cargs), (name, len(fargs), len(cargs))
        for (dummy, d), (how, base, stars) in zip(fargs, cargs):
            assert d is not None, (name, dummy, "dummy argument without a declaration")
            ftype, by_value, is_array = d
            where = (name, dummy, d, (how, base, stars))
            if how == "val":
                assert by_value and ftype == val_kind[base], where
            elif by_value:
                assert ftype == "type(c_ptr)", where  # an opaque address handed through
            elif ftype == "type(c_ptr)":
                assert stars == 2 or base == "void", where  # array of addresses <-> T* const* (or void** out)
            elif base == "void":
                assert stars == 1, where  # untyped buffer: any array by reference
            else:
                assert stars == 1 and ftype == ref_kind.get(base), where
