"""Static check of the Julia `ccall` glue (julia/QaintensorCUDA.jl) against the C header.

Julia is absent from the build image, so the wrapper cannot be executed; what CAN be checked without a
Julia runtime is that every `ccall((:qtn_x, LIB), RetType, (ArgTypes...), ...)` names a function the header declares,
with the same number of parameters, the same return kind and, per parameter, the same machine class (pointer / 32-bit
integer / 64-bit integer / double) -- the transcription errors a first run would otherwise find as a crash."""
import os
import re

from conftest import ROOT

HEADER = open(os.path.join(ROOT, "include", "qaintensor_cuda.h")).read()
JULIA = open(os.path.join(ROOT, "julia", "QaintensorCUDA.jl")).read()


def strip_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


def c_class(t):
    t = t.strip()
    if "*" in t or "[" in t:   # `double cost[2]` decays to a pointer
        return "ptr"
    base = re.sub(r"\b(const|unsigned|signed)\b", "", t).split()
    # drop the parameter name
    words = [w for w in base if w]
    ty = words[0] if words else ""
    return {"int32_t": "i32", "int": "i32", "uint32_t": "i32", "int64_t": "i64", "uint64_t": "i64", "size_t": "i64",
            "double": "f64", "float": "f32", "void": "void"}.get(ty, ty)


def header_protos():
    protos = {}
    for m in re.finditer(r"\b([A-Za-z_][\w \*]*?)\b(qtn_\w+)\s*\(([^;{}]*?)\)\s*;", strip_comments(HEADER)):
        ret, name, params = m.group(1), m.group(2), m.group(3)
        ps = [] if params.strip() in ("", "void") else [c_class(p) for p in params.split(",")]
        protos[name] = ("ptr" if "*" in ret else c_class(ret), ps)
    return protos


def julia_class(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring", "Ptr"):
        return "ptr"
    return {"Cint": "i32", "Int32": "i32", "UInt32": "i32", "Int64": "i64", "UInt64": "i64", "Clonglong": "i64", "Csize_t": "i64",
            "Cdouble": "f64", "Float64": "f64", "Cfloat": "f32", "Cvoid": "void"}.get(t, t)


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def julia_ccalls():
    calls = []
    for m in re.finditer(r"ccall\(\(:(qtn_\w+),\s*LIB\),\s*(\w+),\s*\(", JULIA):
        name, ret = m.group(1), m.group(2)
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(JULIA[i], 0)
            i += 1
        types = [t for t in split_top(JULIA[m.end():i - 1]) if t.strip()]
        # the actual arguments that follow the type tuple, up to the ccall's closing parenthesis
        j, depth = i, 1
        while depth:
            depth += {"(": 1, ")": -1}.get(JULIA[j], 0)
            j += 1
        args = [a for a in split_top(JULIA[i:j - 1].lstrip().lstrip(",")) if a.strip()]
        calls.append((name, julia_class(ret), [julia_class(t) for t in types], len(args)))
    return calls


def test_header_parser_sees_the_abi():
    protos = header_protos()
    assert len(protos) >= 58 and protos["qtn_last_error"] == ("ptr", [])
    assert protos["qtn_svd_trunc"] == ("i32", ["ptr", "i64", "i64", "f64", "i64", "ptr", "ptr", "ptr", "ptr"])


def test_every_ccall_matches_its_prototype():
    protos = header_protos()
    calls = julia_ccalls()
    assert len(calls) >= 18
    for name, ret, types, nargs in calls:
        assert name in protos, "%s is not declared in include/qaintensor_cuda.h" % name
        pret, pparams = protos[name]
        assert ret == pret, "%s: return %s vs header %s" % (name, ret, pret)
        assert types == pparams, "%s: ccall types %s vs header %s" % (name, types, pparams)
        assert nargs == len(types), "%s: %d arguments for %d declared types" % (name, nargs, len(types))


def test_wrapper_covers_the_hot_path_entry_points():
    names = {c[0] for c in julia_ccalls()}
    for need in ("qtn_init", "qtn_contract", "qtn_order_treewidth", "qtn_svd_trunc", "qtn_contract_svd", "qtn_permutedims",
                 "qtn_contract_svd_fold", "qtn_mps_switch_adjacent", "qtn_decompose", "qtn_mpo_from_matrix", "qtn_net_contract"):
        assert need in names


def test_wrapper_messages_are_the_references_own():
    """Every literal `@warn("...")` / `error("...")` text in the wrapper must be a text the reference itself emits
    (src/*.jl), so a user sees the same messages after the switch.  Reads the reference checkout when it is present
    (the build container); the literal list below is the committed copy for machines without it."""
    import glob
    import pytest
    msgs = re.findall(r'(?:@warn|error)\("([^"$]+)"\)', JULIA)
    assert msgs, "no messages found in the wrapper"
    committed = {  # src/network2graph.jl:474, :122, :59; src/svd.jl:9, :16
        "For TensorNetworks with open indices the treewidth algorithm is unlikely to optimize performance",
        "All open indices are disregarded", "Contractions of more than 2 tensors not supported",
        "Error must be positive", "Dimensions of contraction legs do not match"}
    for m in msgs:
        assert m in committed, "wrapper message not among the reference's: %r" % m
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not present: compared with the committed copy only")
    src = "".join(open(f).read() for f in glob.glob(os.path.join(ref, "*.jl")))
    for m in committed:
        assert m in src, "committed message is not in the reference sources: %r" % m
