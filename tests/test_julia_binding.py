"""cloud.jl_b200/julia/SSEB200.jl cannot be executed in this image (no Julia).  What can be checked mechanically is that
the file binds the ABI the header declares: every `ccall((:sym, libsse), Ret, (Args...), ...)` names a declared symbol, with
the declared number of arguments and C-compatible argument / return types, and the struct mirrors SSEConfig / SSEArrays list
the fields of sse_config / sse_arrays in the header's order with matching types (Julia lays isbits structs out like C)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = os.path.join(ROOT, "cloud.jl_b200", "julia", "SSEB200.jl")
HDR = os.path.join(ROOT, "include", "sse_b200.h")


def _header():
    txt = open(HDR).read()
    return re.sub(r"/\*.*?\*/", "", txt, flags=re.S)


def _ctype(t):
    t = t.strip()
    if "*" in t or "[" in t:
        return "ptr"
    t = t.replace("const", "").strip().split()[0]
    return {"int32_t": "i32", "int64_t": "i64", "double": "f64", "void": "void"}[t]


def header_prototypes():
    out = {}
    for ret, name, args in re.findall(r"(int32_t|const char\*)\s+(sse_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", _header(), flags=re.S):
        args = " ".join(args.split())
        al = [] if args in ("void", "") else [_ctype(re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*\s*(\[[0-9]*\])?$", lambda m: m.group(1) or "", a.strip())
                                                     if not a.strip().endswith("*") else a) for a in args.split(",")]
        out[name] = ("ptr" if "char" in ret else "i32", al)
    return out


def _jtype(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Int32": "i32", "Int64": "i64", "Float64": "f64", "Cvoid": "void"}[t]


def _split_top(s):
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        if ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur)
    return [p.strip() for p in parts if p.strip()]


def julia_ccalls():
    txt = open(JL).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(sse_[a-z0-9_]+), libsse\),\s*([A-Za-z0-9{}]+),\s*\(", txt):
        i, depth = m.end(), 1
        while depth:                                       # the argument-type tuple
            depth += {"(": 1, ")": -1}.get(txt[i], 0)
            i += 1
        types = _split_top(txt[m.end():i - 1])
        j, depth, rest = i, 1, ""
        while depth:                                       # the remaining call arguments up to the closing parenthesis of ccall
            ch = txt[j]
            depth += {"(": 1, ")": -1}.get(ch, 0)
            if depth:
                rest += ch
            j += 1
        values = _split_top(rest.lstrip(", \n"))
        calls.append((m.group(1), m.group(2), types, values))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(calls) >= 20
    for name, ret, types, values in calls:
        assert name in protos, f"{name} is not declared in include/sse_b200.h"
        hret, hargs = protos[name]
        assert _jtype(ret) == hret, (name, ret, hret)
        assert len(types) == len(hargs), f"{name}: {len(types)} argument types in the ccall, {len(hargs)} in the header"
        assert len(values) == len(types), f"{name}: {len(values)} values for {len(types)} argument types"
        for k, (jt, ht) in enumerate(zip(types, hargs)):
            assert _jtype(jt) == ht, f"{name} argument {k}: Julia {jt} vs header class {ht}"


def test_the_binding_covers_the_path_and_its_callers():
    bound = {c[0] for c in julia_ccalls()}
    for need in ("sse_create", "sse_destroy", "sse_rhs", "sse_rhs_host", "sse_state_alloc", "sse_state_free", "sse_state_fill",
                 "sse_state_upload", "sse_state_download", "sse_axpby", "sse_step_ck54", "sse_functionals", "sse_synchronize",
                 "sse_comm_init_all", "sse_halo_plan", "sse_rhs_multi", "sse_step_ck54_multi", "sse_last_error_string"):
        assert need in bound, need


def _struct_fields_header(name):
    m = re.search(r"typedef struct " + name + r" \{(.*?)\} " + name + ";", _header(), flags=re.S)
    fields = []
    for decl in m.group(1).split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        base, names = re.match(r"((?:const\s+)?[a-z0-9_]+\s*\*?)\s*(.*)", decl).groups()
        for n in names.split(","):
            n = n.strip()
            arr = re.search(r"\[(\d+)\]", n)
            n0 = re.sub(r"\[.*\]", "", n).strip()
            elt = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64"}[base.replace("const", "").replace("*", "").strip()]
            t = f"Ptr{{{elt}}}" if "*" in base else elt
            fields.append((n0, f"NTuple{{3, {t}}}" if arr else t))
    return fields


def _struct_fields_julia(name):
    m = re.search(r"struct " + name + r"\n(.*?)\nend", open(JL).read(), flags=re.S)
    fields = []
    for part in re.split(r"[;\n]", m.group(1)):
        part = part.strip()
        if part:
            n, t = part.split("::")
            fields.append((n.strip(), t.strip()))
    return fields


def test_struct_mirrors_follow_the_header():
    assert _struct_fields_julia("SSEConfig") == _struct_fields_header("sse_config")
    assert _struct_fields_julia("SSEArrays") == _struct_fields_header("sse_arrays")
