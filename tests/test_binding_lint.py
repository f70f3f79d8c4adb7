"""Signature lint of the two host bindings against include/mktfhe_b200.h.

julia/TFHE_B200.jl cannot be executed in this image (no Julia), so its `ccall`s are parsed and compared with the C declarations:
symbol exists, return type, number of arguments, and the C type each Julia type maps to.  The ctypes table of
torus-fhe_b200/_cabi.py gets the same treatment, so that the binding the tests drive and the binding a Julia host uses cannot drift
apart from the header unnoticed.  INTEGRATION.md's snippets are checked for symbols that do not exist.
"""
import ctypes as C
import os
import re

from conftest import ROOT


def c_declarations():
    """name -> (return type, [parameter types]) with names and `const` stripped, pointers kept."""
    text = open(os.path.join(ROOT, "include", "mktfhe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for ret, name, args in re.findall(r"^\s*((?:const\s+)?[a-z_0-9]+\s*\**)\s*(mktfhe_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text, flags=re.M):
        params = []
        for a in ([] if args.strip() == "void" else args.split(",")):
            a = a.replace("const", " ").strip()
            m = re.match(r"([a-z_0-9]+)\s*(\**)\s*[A-Za-z_0-9]*$", a)
            assert m, (name, a)
            params.append(m.group(1) + m.group(2))
        decls[name] = (re.sub(r"\s+", "", ret.replace("const", "")), params)
    return decls


# Julia ccall type -> the C types it may stand for
JULIA_TO_C = {
    "Ptr{Cvoid}": {"mktfhe_ctx*", "void*"},
    "Ref{Ptr{Cvoid}}": {"mktfhe_ctx**", "void**"},
    "Ref{CParams}": {"mktfhe_params*"},
    "Cint": {"int"},
    "Ptr{Cint}": {"int*"},
    "Cvoid": {"void"},
    "Csize_t": {"size_t"},
    "Int64": {"int64_t"},
    "Int32": {"int32_t"},
    "UInt64": {"uint64_t"},
    "Ptr{Int32}": {"int32_t*"},
    "Ptr{Int64}": {"int64_t*"},
    "Cstring": {"char*"},
}


def split_top(s):
    """split on commas that are not inside (), {} or []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls(fname="TFHE_B200.jl"):
    text = open(os.path.join(ROOT, "julia", fname)).read()
    text = "\n".join(line.split("#")[0] if not line.lstrip().startswith("#") else "" for line in text.splitlines())
    calls = []
    for m in re.finditer(r"ccall\(\(:(mktfhe_[a-z0-9_]+),\s*LIB\)\s*,", text):
        # take the balanced argument list of this ccall
        i = m.start() + len("ccall")
        depth, j = 0, i
        while True:
            depth += text[j] == "("
            depth -= text[j] == ")"
            j += 1
            if depth == 0:
                break
        parts = split_top(text[i + 1:j - 1])
        ret, types, args = parts[1], parts[2], parts[3:]
        assert types.startswith("(") and types.endswith(")"), (m.group(1), types)
        calls.append((m.group(1), ret, split_top(types[1:-1]), args))
    return calls


def test_julia_ccalls_match_the_header():
    decls = c_declarations()
    calls = julia_ccalls()
    assert len(calls) >= 12
    for name, ret, types, args in calls:
        assert name in decls, f"{name} is not declared in include/mktfhe_b200.h"
        cret, cparams = decls[name]
        assert cret in JULIA_TO_C[ret], (name, ret, cret)
        assert len(types) == len(cparams), (name, types, cparams)
        assert len(args) == len(types), f"{name}: {len(args)} values for {len(types)} declared types"
        for jt, ct in zip(types, cparams):
            assert jt in JULIA_TO_C, (name, jt)
            assert ct in JULIA_TO_C[jt], f"{name}: Julia {jt} bound to C {ct}"
    # every entry point of the hot path and of key loading is bound by the shim
    bound = {c[0] for c in calls}
    for must in ("mktfhe_create_multi", "mktfhe_device_count", "mktfhe_destroy", "mktfhe_last_error", "mktfhe_load_bsk", "mktfhe_load_ksk", "mktfhe_finalize_keys",
                 "mktfhe_bootstrap_batch", "mktfhe_gate_batch", "mktfhe_gate_batch_mixed", "mktfhe_blind_rotate_batch", "mktfhe_keyswitch_batch"):
        assert must in bound, must


def test_single_key_julia_shim_matches_the_header():
    """julia/TFHE1_B200.jl (single-key gate API over the same library): every ccall against the C declarations, the CParams
    struct, balanced blocks, every exported name defined, and the same gate table as the Python twin / the oracle (gates.jl:16-142)."""
    decls = c_declarations()
    calls = julia_ccalls("TFHE1_B200.jl")
    assert len(calls) >= 9
    for name, ret, types, args in calls:
        assert name in decls, name
        cret, cparams = decls[name]
        assert cret in JULIA_TO_C[ret], (name, ret, cret)
        assert len(types) == len(cparams) == len(args), (name, types, cparams, args)
        for jt, ct in zip(types, cparams):
            assert ct in JULIA_TO_C[jt], f"{name}: Julia {jt} bound to C {ct}"
    bound = {c[0] for c in calls}
    for must in ("mktfhe_create_multi", "mktfhe_load_bsk", "mktfhe_load_ksk", "mktfhe_finalize_keys", "mktfhe_affine_bootstrap_batch",
                 "mktfhe_blind_rotate_batch", "mktfhe_keyswitch_batch", "mktfhe_bootstrap_batch", "mktfhe_destroy"):
        assert must in bound, must
    raw = open(os.path.join(ROOT, "julia", "TFHE1_B200.jl")).read()
    text = _julia_balance(raw)
    body = re.search(r"struct CParams.*?\n(.*?)\nend", raw, flags=re.S).group(1)
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mktfhe_b200.h")).read(), flags=re.S)
    cfields = re.findall(r"int32_t\s+([A-Za-z_]+);", re.search(r"typedef struct \{(.*?)\} mktfhe_params;", header, flags=re.S).group(1))
    assert re.findall(r"([A-Za-z_]+)::Int32", body) == cfields
    exported = re.search(r"\nexport (.*?)\n\n", text + "\n\n", flags=re.S).group(1)
    for name in re.findall(r"[A-Za-z_][A-Za-z_0-9!]*", exported):
        assert re.search(rf"(function |struct |^|\(:){re.escape(name)}(?![A-Za-z_0-9!])", text, flags=re.M), f"exported {name} is not defined"
    from oracle import tfhe1_oracle as O1
    table = dict((m.group(1).upper(), tuple(int(v) for v in m.group(2, 3, 4, 5)))
                 for m in re.finditer(r"\(:gate_([a-z]+), (-?\d+), (\d+), (-?\d+), (-?\d+)\)", raw))
    assert table == O1.GATE_LINEAR


def test_ccs_julia_shim_matches_the_header():
    """julia/TFHE_CCS_B200.jl (the CCS multi-key gate API over the same library): every ccall against the C declarations, the CParams struct,
    balanced blocks, every exported name defined, and the same pseudo-party numbering of the hybrid product's key elements as the Python
    twin (tfhe_ccs.build_elements)."""
    decls = c_declarations()
    calls = julia_ccalls("TFHE_CCS_B200.jl")
    assert len(calls) >= 10
    for name, ret, types, args in calls:
        assert name in decls, name
        cret, cparams = decls[name]
        assert cret in JULIA_TO_C[ret], (name, ret, cret)
        assert len(types) == len(cparams) == len(args), (name, types, cparams, args)
        for jt, ct in zip(types, cparams):
            assert ct in JULIA_TO_C[jt], f"{name}: Julia {jt} bound to C {ct}"
    bound = {c[0] for c in calls}
    for must in ("mktfhe_create", "mktfhe_load_bsk", "mktfhe_load_ksk", "mktfhe_mark_keys_received", "mktfhe_finalize_keys",
                 "mktfhe_ccs_blind_rotate_batch", "mktfhe_mk_keyswitch_batch", "mktfhe_destroy", "mktfhe_last_error"):
        assert must in bound, must
    raw = open(os.path.join(ROOT, "julia", "TFHE_CCS_B200.jl")).read()
    text = _julia_balance(raw)
    body = re.search(r"struct CParams.*?\n(.*?)\nend", raw, flags=re.S).group(1)
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mktfhe_b200.h")).read(), flags=re.S)
    cfields = re.findall(r"int32_t\s+([A-Za-z_]+);", re.search(r"typedef struct \{(.*?)\} mktfhe_params;", header, flags=re.S).group(1))
    assert re.findall(r"([A-Za-z_]+)::Int32", body) == cfields
    exported = re.search(r"\nexport (.*?)\n\n", text + "\n\n", flags=re.S).group(1)
    for name in re.findall(r"[A-Za-z_][A-Za-z_0-9!]*", exported):
        assert re.search(rf"(function |struct |^){re.escape(name)}(?![A-Za-z_0-9!])", text, flags=re.M), f"exported {name} is not defined"
    # pseudo-party numbering: round 1 of polynomial i in party p's steps = p (k + 1) + i, round 2 = k (k + 1) + p (0-based), as in the twin
    assert "(party - 1) * (k + 1) + i" in raw and "k * (k + 1) + party - 1" in raw
    twin = open(os.path.join(ROOT, "torus-fhe_b200", "tfhe_ccs.py")).read()
    assert "elems[pi * (k + 1) + i" in twin and "elems[k * (k + 1) + pi" in twin
    assert re.search(r"MKTFHE_FLAG_TORUS32\s+1", header) and "FLAG_TORUS32 = Int32(1)" in raw


def test_julia_cparams_struct_matches_mktfhe_params():
    text = open(os.path.join(ROOT, "julia", "TFHE_B200.jl")).read()
    body = re.search(r"struct CParams.*?\n(.*?)\nend", text, flags=re.S).group(1)
    fields = re.findall(r"([A-Za-z_]+)::Int32", body)
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "mktfhe_b200.h")).read(), flags=re.S)
    cfields = re.findall(r"int32_t\s+([A-Za-z_]+);", re.search(r"typedef struct \{(.*?)\} mktfhe_params;", header, flags=re.S).group(1))
    assert fields == cfields
    import torus_fhe_b200 as T
    assert [f[0] for f in T._cabi.CParams._fields_] == cfields and C.sizeof(T._cabi.CParams) == 4 * len(cfields)
    # gate ids
    ids = dict(re.findall(r"#define MKTFHE_GATE_([A-Z0-9]+) (\d+)", header))
    jl = re.search(r"const GATE_NAND, GATE_OR, GATE_AND, GATE_XOR, GATE_AND3 = (.*)", text).group(1)
    assert [int(x) for x in re.findall(r"Cint\((\d+)\)", jl)] == [int(ids[k]) for k in ("NAND", "OR", "AND", "XOR", "AND3")]
    assert (T._cabi.GATE_NAND, T._cabi.GATE_OR, T._cabi.GATE_AND, T._cabi.GATE_XOR, T._cabi.GATE_AND3) == tuple(
        int(ids[k]) for k in ("NAND", "OR", "AND", "XOR", "AND3"))


def ctypes_to_c(t):
    """C types a ctypes type may stand for (on LP64 c_uint64, c_size_t and c_ulong are one class, so match by size and sign)."""
    if t is None:
        return {"void"}
    if t is C.c_void_p:
        return {"mktfhe_ctx*", "void*", "int32_t*", "int64_t*"}
    if t is C.c_char_p:
        return {"char*"}
    out = set()
    if t is C.c_double:
        return {"double"}
    for name, ct in (("int", C.c_int), ("int64_t", C.c_int64), ("uint64_t", C.c_uint64), ("size_t", C.c_size_t), ("int32_t", C.c_int32)):
        if C.sizeof(ct) == C.sizeof(t) and (ct(-1).value < 0) == (t(-1).value < 0):
            out.add(name)
    return out


def test_ctypes_table_matches_the_header():
    import torus_fhe_b200 as T
    L = T._cabi.lib()
    decls = c_declarations()
    assert set(decls) == set(T._cabi.EXPORTS)
    for name, (cret, cparams) in decls.items():
        f = getattr(L, name)
        assert cret in ctypes_to_c(f.restype), (name, f.restype, cret)
        assert len(f.argtypes) == len(cparams), (name, f.argtypes, cparams)
        for at, ct in zip(f.argtypes, cparams):
            if not hasattr(at, "_type_") or at in (C.c_void_p, C.c_char_p) or isinstance(at._type_, str):
                assert ct in ctypes_to_c(at), (name, at, ct)
            else:                                       # POINTER(x): one more level of indirection than x
                assert ct.endswith("*"), (name, at, ct)
                inner = at._type_
                base = ct[:-1]
                ok = {T._cabi.CParams: {"mktfhe_params"}, C.c_void_p: {"mktfhe_ctx*", "void*"}, C.c_size_t: {"size_t"}, C.c_float: {"float"},
                      C.c_double: {"double"}, C.c_int: {"int"}}[inner]
                assert base in ok, (name, at, ct)


def test_integration_doc_names_only_existing_symbols():
    decls = c_declarations()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    used = set(re.findall(r"\b(mktfhe_[a-z0-9_]+)\b", text))
    used -= {"mktfhe_b200", "mktfhe_params", "mktfhe_ctx", "mktfhe_parameters_2party_3gen"}
    used = {u for u in used if not u.startswith("mktfhe_parameters_")}
    assert used and used <= set(decls), used - set(decls)


def _julia_balance(text):
    # strip docstrings / strings / comments
    text = re.sub(r'"""(.*?)"""', '""', text, flags=re.S)
    text = re.sub(r'"(?:\\.|[^"\\\n])*"', '""', text)
    assert text.count('"') % 2 == 0
    text = "\n".join(line.split("#")[0] for line in text.splitlines())
    depth, blocks, stack = 0, 0, []
    pairs = {")": "(", "]": "[", "}": "{"}
    for tok in re.findall(r"[A-Za-z_][A-Za-z_0-9!]*|[()\[\]{}]", text):
        if tok in "([{":
            stack.append(tok)
        elif tok in ")]}":
            assert stack and stack.pop() == pairs[tok], "unbalanced bracket"
        elif not stack:
            if tok in ("module", "struct", "function", "for", "if", "while", "do", "begin", "let", "try", "quote"):
                blocks += 1
            elif tok == "end":
                blocks -= 1
                assert blocks >= 0, "`end` without an opener"
    assert not stack and blocks == 0, (stack, blocks)
    return text


def test_julia_files_blocks_and_brackets_balance():
    """No Julia here to parse the files: at least every block opener outside brackets has its `end`, brackets nest, strings close,
    every exported name of the shim is defined, and the fixture script only calls writer functions the shim defines."""
    text = _julia_balance(open(os.path.join(ROOT, "julia", "TFHE_B200.jl")).read())
    dump = _julia_balance(open(os.path.join(ROOT, "julia", "dump_fixture.jl")).read())
    for name in set(re.findall(r"TFHE_B200\.([A-Za-z_0-9]+)", dump)):
        assert re.search(rf"(function |struct ){name}\b", text), f"dump_fixture.jl uses TFHE_B200.{name}, which the shim does not define"
    exported = re.search(r"\nexport (.*?)\n\n", text + "\n\n", flags=re.S).group(1)
    for name in re.findall(r"[A-Za-z_][A-Za-z_0-9]*", exported):
        assert re.search(rf"(function |struct |^|\(:){name}\b", text, flags=re.M), f"exported {name} is not defined"
