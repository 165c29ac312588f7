"""CPU-side checks: the C-ABI library loads and exports every symbol include/ifadv.h declares (no compute calls),
the host logic (sweep-order rotation, masks), and that the product fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib_path():
    return os.path.join(ROOT, "interfaceadvection.jl_b200", "libifadv_b200.so")


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ifadv.h")).read()
    names = sorted(set(re.findall(r"\b(ifadv_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 15
    if not os.path.exists(_lib_path()):
        pytest.fail("libifadv_b200.so is not built: run __graft_entry__.build()")
    L = ctypes.CDLL(_lib_path())
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ifadv.h but not exported"
    L.ifadv_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.ifadv_version()


def test_no_cpu_fallback():
    import torch

    import interfaceadvection.jl_b200 as ia

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ia.IfadvError):
        ia.Context((10, 10), "float64", 0)
    with pytest.raises(ia.IfadvError):
        ia.BCf(torch.zeros(4, 4))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the product package or include/ may reference it."""
    pkg = os.path.join(ROOT, "interfaceadvection.jl_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".jl")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, os.path.join(dp, fn)


def test_sweep_order_rotation():
    from interfaceadvection.jl_b200.api import _dirO

    class F:
        pass

    f = F()
    f.dt = [0.25]
    assert _dirO(f, 3) == (3, 1, 2) and _dirO(f, 2) == (1, 2)
    f.dt = [0.25, 0.1]
    assert _dirO(f, 3) == (1, 2, 3) and _dirO(f, 2) == (2, 1)
    f.dt = [0.25, 0.1, 0.1]
    assert _dirO(f, 3) == (2, 3, 1)


def test_perdir_mask():
    from interfaceadvection.jl_b200._lib import perdir_mask

    assert perdir_mask(()) == 0 and perdir_mask((1,)) == 1 and perdir_mask((2, 3)) == 6 and perdir_mask((1, 2, 3)) == 7


def test_bench_byte_accounting_matches_survey():
    """SURVEY §8d figures the roofline fractions in bench.py are built on: 53 / 105 B per cell and CMOM sweep, 376 / 744 B per CMOM
    step, 52 / 100 B per pure-VOF step (3-D), 67 B for the 2-D Float64 VOF step."""
    import importlib.util
    import os

    spec = importlib.util.spec_from_file_location("bench", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    assert b.algorithmic_bytes_per_cell_sweep(3, 4) == 53 and b.algorithmic_bytes_per_cell_sweep(3, 8) == 105
    assert b.algorithmic_bytes_per_cell_step(3, 4) == 376 and b.algorithmic_bytes_per_cell_step(3, 8) == 744
    assert b.vof_bytes_per_cell_step(3, 4) == 52 and b.vof_bytes_per_cell_step(3, 8) == 100 and b.vof_bytes_per_cell_step(2, 8) == 67
    assert set(b.WORKLOADS) >= {"C4_bubble_512_f32", "C3_dambreak_512x256x256_f32", "C2_enright_256_f32", "C1_zalesak_128_f64"}


def _c_kind(arg: str) -> str:
    arg = arg.strip()
    if "*" in arg or "[" in arg:
        return "ptr"
    if arg.startswith("double"):
        return "double"
    if arg.startswith("unsigned"):
        return "uint"
    if arg.startswith("int64_t"):
        return "int64"
    if arg.startswith("int"):
        return "int"
    raise AssertionError(arg)


def _jl_kind(t: str) -> str:
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Cdouble": "double", "Cuint": "uint", "Cint": "int", "Int64": "int64"}[t]


def test_julia_ccall_signatures_match_header():
    """The Julia glue cannot be executed here (no Julia in the image): at least hold every `ccall` in it against the prototype in
    include/ifadv.h -- same symbol, same number of arguments, same C kind (pointer / double / int / unsigned) position by position,
    and the report struct field for field."""
    hdr = open(os.path.join(ROOT, "include", "ifadv.h")).read()
    hdr_nc = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(ifadv_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr_nc, flags=re.S):
        args = [a for a in re.split(r",", m.group(2).replace("\n", " ")) if a.strip() and a.strip() != "void"]
        protos[m.group(1)] = [_c_kind(a) for a in args]
    jl = open(os.path.join(ROOT, "interfaceadvection.jl_b200", "julia", "IntfAdvB200Ext.jl")).read()
    calls = re.findall(r"ccall\(\(:(ifadv_[a-z0-9_]+), LIB\),\s*\w+,\s*\((.*?)\),\s*\n?\s*ctx", jl, flags=re.S)
    seen = set()
    for name, types in calls:
        # split the Julia type tuple at top-level commas
        parts, depth, cur = [], 0, ""
        for ch in types.replace("\n", " "):
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur); cur = ""
            else:
                cur += ch
        if cur.strip():
            parts.append(cur)
        kinds = [_jl_kind(p) for p in parts]
        assert name in protos, name
        assert kinds == protos[name], (name, kinds, protos[name])
        seen.add(name)
    assert {"ifadv_create", "ifadv_advect_vof", "ifadv_advect_vof_rhouu", "ifadv_u2rhou_advect_vof_rhouu", "ifadv_u2rhou", "ifadv_rhou2u",
            "ifadv_mpcfl", "ifadv_last_error", "ifadv_poisson_update", "ifadv_psolver", "ifadv_myproject", "ifadv_ml_create", "ifadv_ml_destroy",
            "ifadv_ml_update", "ifadv_ml_solver", "ifadv_ml_myproject"} <= seen
    # report struct: field order and C types
    c_fields = re.search(r"typedef struct \{(.*?)\} ifadv_report;", hdr_nc, flags=re.S).group(1)
    c_names = re.findall(r"\b(maxf|minf|argmax|argmin|dir|status|div_u0|div_u)\b", c_fields)
    j_fields = re.search(r"struct IfadvReport\n(.*?)\nend", jl, flags=re.S).group(1)
    j_names = re.findall(r"\b(maxf|minf|argmax|argmin|dir|status|div_u0|div_u)::", j_fields)
    assert c_names == j_names == ["maxf", "minf", "argmax", "argmin", "dir", "status", "div_u0", "div_u"]
    import ctypes as C
    from interfaceadvection.jl_b200._lib import Report
    assert [f[0] for f in Report._fields_] == c_names and C.sizeof(Report) == 8 * 2 + 8 * 6 + 4 * 2 + 8 * 2
