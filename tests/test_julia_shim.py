"""The Julia ccall shim (julia/UnsflowCUDA.jl) cannot be executed here (no julia binary), so its
FFI surface is checked statically against include/unsflow.h: every ccall names an exported symbol,
passes the right number of arguments with C-compatible types, and the struct mirrors list the same
fields, in the same order and of the same width, as the C structs."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = open(os.path.join(ROOT, "julia", "UnsflowCUDA.jl")).read()
HDR = open(os.path.join(ROOT, "include", "unsflow.h")).read()


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _c_prototypes():
    text = re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|const char \*|void)\s*(unsflow_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        args = " ".join(m.group(3).split())
        protos[m.group(2)] = [] if args in ("void", "") else [a.strip() for a in args.split(",")]
    return protos


def _c_kind(arg):
    """C parameter -> coarse kind"""
    a = arg.replace("const ", "").strip()
    if "**" in a:
        return "ptrptr"
    if "*" in a:
        base = a.split("*")[0].strip()
        return {"double": "ptr_f64", "float": "ptr_f32", "int64_t": "ptr_i64", "int": "ptr_i32", "int32_t": "ptr_i32",
                "void": "ptr_void", "char": "ptr_char"}.get(base, "ptr_struct:" + base)
    base = a.rsplit(" ", 1)[0].strip()
    return {"double": "f64", "float": "f32", "int64_t": "i64", "int": "i32", "int32_t": "i32"}[base]


def _jl_kind(t):
    t = t.strip()
    table = {"Int64": "i64", "Cint": "i32", "Int32": "i32", "Cdouble": "f64", "Float64": "f64", "Cfloat": "f32",
             "Ptr{Cdouble}": "ptr_f64", "Ptr{Float64}": "ptr_f64", "Ptr{Int64}": "ptr_i64", "Ptr{Cint}": "ptr_i32",
             "Ptr{Cvoid}": "ptr_void", "Ptr{Ptr{Cvoid}}": "ptrptr", "Ptr{UInt8}": "ptr_void", "Ptr{Cfloat}": "ptr_f32",
             "Ptr{CSurf}": "ptr_struct:unsflow_surf", "Ptr{CField}": "ptr_struct:unsflow_field", "Ptr{COpts}": "ptr_struct:unsflow_opts"}
    return table[t]


def _compatible(c, j):
    if c == j:
        return True
    # opaque handles are passed as Ptr{Cvoid}
    if c.startswith("ptr_struct:unsflow_ldvm") or c.startswith("ptr_struct:unsflow_wake"):
        return j == "ptr_void"
    return False


def test_every_ccall_matches_the_header():
    protos = _c_prototypes()
    calls = list(re.finditer(r"ccall\(\(:(\w+), LIB\),\s*(\w+),\s*\(", JL))
    assert len(calls) >= 20
    seen = set()
    for m in calls:
        name, ret = m.group(1), m.group(2)
        assert name in protos, f"{name} is not declared in include/unsflow.h"
        seen.add(name)
        # argument-type tuple: balanced parentheses from the '(' that ends the match
        i = m.end() - 1
        depth, j = 0, i
        while True:
            depth += JL[j] == "("
            depth -= JL[j] == ")"
            if depth == 0:
                break
            j += 1
        types = [t for t in _split_top(JL[i + 1:j]) if t]
        cargs = protos[name]
        assert len(types) == len(cargs), f"{name}: {len(types)} ccall argument types vs {len(cargs)} C parameters"
        for t, c in zip(types, cargs):
            assert _compatible(_c_kind(c), _jl_kind(t)), f"{name}: Julia {t} vs C '{c}'"
        assert ret == ("Cstring" if name == "unsflow_last_error" else "Cint")
        # values passed: same count as types
        rest = JL[j + 1:]
        depth, k = 1, 0       # we are inside ccall( ... ; find its closing parenthesis
        while depth:
            depth += rest[k] in "([{"
            depth -= rest[k] in ")]}"
            k += 1
        vals = [v for v in _split_top(rest[:k - 1].lstrip(", \n")) if v]
        assert len(vals) == len(types), f"{name}: {len(vals)} values for {len(types)} argument types"
    # the shim covers the whole reference-facing surface
    for need in ("unsflow_ind_vel", "unsflow_mutual_ind", "unsflow_wakeroll", "unsflow_update_indbound", "unsflow_ldvm_create",
                 "unsflow_ldvm_set_kinematics", "unsflow_ldvm_set_2dof", "unsflow_ldvm_get_2dof", "unsflow_ldvm_run",
                 "unsflow_ldvm_counts", "unsflow_ldvm_get_state", "unsflow_ldvm_destroy", "unsflow_ldvm_multi_create",
                 "unsflow_ldvm_multi_get_state", "unsflow_lve_create", "unsflow_lve_get_circ", "unsflow_last_error"):
        assert need in seen, f"shim never calls {need}"


def _c_struct_fields(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HDR, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        base, names = decl.rsplit(" ", 1)[0], None
        m = re.match(r"(unsflow_vorts|double|int32_t|int64_t)\s+(.*)", decl)
        base, names = m.group(1), m.group(2)
        for n in names.split(","):
            n = n.strip()
            arr = re.match(r"(\w+)\[(\d+)\]", n)
            if arr:
                fields.append((arr.group(1), f"{base}[{arr.group(2)}]"))
            elif n.startswith("*"):
                fields.append((n[1:], base + "*"))
            else:
                fields.append((n, base))
    return fields


def _jl_struct_fields(name):
    body = re.search(r"struct %s\b[^\n]*\n(.*?)\nend" % name, JL, flags=re.S).group(1)
    fields = []
    for part in re.split(r"[;\n]", body):
        part = part.split("#")[0].strip()
        if part:
            n, t = part.split("::")
            fields.append((n.strip(), t.strip()))
    return fields


def test_struct_mirrors_match_field_by_field():
    jl2c = {"Int64": "int64_t", "Int32": "int32_t", "Cdouble": "double", "Ptr{Cdouble}": "double*", "NTuple{6,Cdouble}": "double[6]",
            "CVorts": "unsflow_vorts"}
    for jname, cname in (("CVorts", "unsflow_vorts"), ("CSurf", "unsflow_surf"), ("CField", "unsflow_field"), ("COpts", "unsflow_opts")):
        cf, jf = _c_struct_fields(cname), _jl_struct_fields(jname)
        assert [n for n, _ in cf] == [n for n, _ in jf], f"{jname}: field names/order differ from {cname}"
        for (n, ct), (_, jt) in zip(cf, jf):
            assert jl2c[jt] == ct, f"{jname}.{n}: Julia {jt} vs C {ct}"
