"""CPU: the Julia glue (reducedbasismethods.jl_b200/julia/...) cannot be executed in the build image (no Julia), so its
`ccall` signatures are checked statically against the C prototypes of include/rbm_b200.h: every called symbol exists,
has the same number of parameters, and every parameter has a compatible Julia C-type.  (A wrong arity or an Int64
passed where the ABI takes an int corrupts the call silently; this is the class of defect a write-only binding hides.)"""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rbm_b200.h")
GLUE = os.path.join(ROOT, "reducedbasismethods.jl_b200", "julia", "ReducedBasisMethodsB200", "src", "ReducedBasisMethodsB200.jl")

# C parameter type (normalised) -> acceptable Julia ccall argument types
COMPAT = {
    "int": {"Cint"},
    "int64_t": {"Int64"},
    "uint64_t": {"UInt64"},
    "double": {"Cdouble"},
    "double*": {"Ptr{Cdouble}", "Ref{Cdouble}"},
    "int*": {"Ptr{Cint}", "Ref{Cint}"},
    "int32_t*": {"Ptr{Int32}", "Ptr{Cint}"},
    "int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
    "void*": {"Ptr{Cvoid}", "Ptr{UInt8}", "Ptr{Cdouble}"},
    "handle*": {"Ptr{Cvoid}"},                 # rbm_ctx*, rbm_pic*, rbm_reduced*, rbm_evd*
    "handle**": {"Ref{Ptr{Cvoid}}"},
    "void**": {"Ref{Ptr{Cvoid}}"},
    "double**": {"Ptr{Ptr{Cdouble}}"},         # const double *const *blocks
}
RET = {"int": "Cint", "void": "Cvoid", "const char*": "Cstring", "int64_t": "Int64"}


def c_prototypes():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(int|void|int64_t|const char \*)\s*\b(rbm_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        ret, name, args = m.group(1).replace(" ", "").replace("constchar*", "const char*"), m.group(2), m.group(3)
        params = []
        if args.strip() not in ("", "void"):
            for a in args.split(","):
                a = re.sub(r"\bconst\b", "", a).strip()
                stars = a.count("*")
                base = re.sub(r"[*]", " ", a).split()
                typ = base[0] if len(base) > 1 or stars else a
                if typ.startswith("rbm_"):
                    typ = "handle"
                params.append(typ + "*" * stars)
        protos[name] = (ret, params)
    return protos


def split_top(s):
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


def julia_ccalls():
    src = open(GLUE).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(\w+),\s*librbm\),\s*(\w+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:                      # the argument-type tuple
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        types = split_top(src[m.end():i - 1])
        j, depth = i, 1                   # the rest of the ccall: the actual arguments
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        rest = src[i:j - 1].lstrip(", \n")
        calls.append((m.group(1), m.group(2), types, split_top(rest), src.count("\n", 0, m.start()) + 1))
    return calls


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    assert len(protos) >= 40, "header parse lost prototypes"
    calls = julia_ccalls()
    assert len(calls) >= 25, "glue parse lost ccalls"
    for name, ret, types, args, line in calls:
        assert name in protos, f"line {line}: {name} is not declared in rbm_b200.h"
        cret, cparams = protos[name]
        assert RET[cret] == ret, f"line {line}: {name} returns {cret}, glue says {ret}"
        assert len(types) == len(cparams), f"line {line}: {name} takes {len(cparams)} parameters, glue passes {len(types)} types"
        assert len(args) == len(types), f"line {line}: {name}: {len(types)} argument types but {len(args)} arguments"
        for k, (jt, ct) in enumerate(zip(types, cparams)):
            assert jt in COMPAT[ct], f"line {line}: {name} parameter {k + 1} is `{ct}`, glue passes `{jt}`"


def test_glue_covers_the_entry_points_a_julia_host_needs():
    called = {c[0] for c in julia_ccalls()}
    for need in ("rbm_init", "rbm_finalize", "rbm_last_error", "rbm_pic_create", "rbm_pic_destroy", "rbm_pic_set_particles",
                 "rbm_pic_run", "rbm_field_update", "rbm_field_eval", "rbm_field_energy", "rbm_field_coefficients",
                 "rbm_pod_factor", "rbm_pod_basis", "rbm_evd_destroy", "rbm_deim", "rbm_reduced_create", "rbm_reduced_run",
                 "rbm_reduced_destroy", "rbm_host_register", "rbm_host_unregister", "rbm_comm_init", "rbm_comm_unique_id",
                 "rbm_sample_bump_on_tail", "rbm_growth_rate"):
        assert need in called, need
