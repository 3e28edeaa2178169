"""Julia cannot run in this image, so the `ccall` shim (julia/QuadrupedLandingB200.jl) is checked statically: every
ccall's symbol, arity, argument types and return type against the prototypes of include/qlnlp.h, and the Julia mirror
structs against the C structs field by field.  The library itself must export every symbol the shim binds."""
import os
import re

import quadruped_landing_b200 as ql

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JL = open(os.path.join(ROOT, "julia", "QuadrupedLandingB200.jl")).read()
HDR = open(os.path.join(ROOT, "include", "qlnlp.h")).read()

# what a C parameter type may be bound to in a ccall signature
ACCEPT = {
    "qlnlp_handle": {"Ptr{Cvoid}"},
    "qlnlp_handle*": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "const qlnlp_problem_desc*": {"Ref{QlProblemDesc}", "Ptr{QlProblemDesc}"},
    "const qlnlp_batch_io*": {"Ref{QlBatchIO}", "Ptr{QlBatchIO}"},
    "int": {"Cint"},
    "int64_t": {"Int64"},
    "double": {"Cdouble"},
    "const int*": {"Ptr{Cint}", "Ref{Cint}"},
    "int*": {"Ptr{Cint}", "Ref{Cint}"},
    "const double*": {"Ptr{Cdouble}", "Ref{Cdouble}"},
    "double*": {"Ptr{Cdouble}", "Ref{Cdouble}"},
    "int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
    "const int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
    "void*": {"Ptr{Cvoid}"},
    "void**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "const char*": {"Cstring", "Ptr{UInt8}"},
    "void": set(),
}
RET = {"int": "Cint", "const char*": "Cstring"}


def c_prototypes():
    body = re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)
    protos = {}
    for ret, name, args in re.findall(r"^\s*(int|const char\*)\s+(qlnlp_\w+)\s*\(([^)]*)\)\s*;", body, flags=re.M):
        params = []
        for a in [x.strip() for x in args.split(",") if x.strip()]:
            if a == "void":
                continue
            a = re.sub(r"\[\d*\]", "*", a)                            # int64_t info[5] -> int64_t info*
            m = re.match(r"^(.*?)(\w+)\s*(\**)$", a)                 # type name [*]
            typ = (m.group(1).strip() + m.group(3)).replace(" *", "*")
            typ = typ.replace("void* const*", "void**")
            params.append(typ)
        protos[name] = (ret, params)
    return protos


def split_types(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "{":
            depth += 1
        if ch == "}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_every_ccall_matches_the_header():
    protos = c_prototypes()
    calls = re.findall(r"ccall\(\(:(\w+),\s*LIBQLNLP\),\s*(\w+),\s*\(([^)]*)\)", JL)
    assert len(calls) >= 18
    L = ql.load_library()
    for sym, ret, argt in calls:
        assert sym in protos, f"{sym} is not declared in include/qlnlp.h"
        assert hasattr(L, sym), f"{sym} is not exported by libqlnlp.so"
        cret, cparams = protos[sym]
        assert RET[cret] == ret, (sym, ret, cret)
        jl = split_types(argt)
        assert len(jl) == len(cparams), (sym, jl, cparams)
        for j, c in zip(jl, cparams):
            assert j in ACCEPT[c], f"{sym}: Julia {j} bound to C parameter {c}"
    bound = {c[0] for c in calls}
    # the seven MOI callbacks + structure + lifecycle are all bound
    for need in ("qlnlp_create", "qlnlp_create_multi", "qlnlp_destroy", "qlnlp_dims", "qlnlp_jacobian_structure",
                 "qlnlp_eval_objective", "qlnlp_eval_objective_gradient", "qlnlp_eval_constraint",
                 "qlnlp_eval_constraint_jacobian", "qlnlp_eval_all", "qlnlp_eval_batch_host", "qlnlp_last_error",
                 "qlnlp_host_output_register", "qlnlp_constraint_bounds", "qlnlp_variable_bounds"):
        assert need in bound, need


def _c_struct(name):
    m = re.search(r"typedef struct \{([^{}]*)\}\s*" + name + r"\s*;", HDR, flags=re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    fields = []
    for decl in [" ".join(d.split()) for d in body.split(";") if d.strip()]:
        m2 = re.match(r"^(const\s+)?(\w+)\s*(\*?)\s*(.*)$", decl)
        typ = m2.group(2) + m2.group(3)
        for nm in [x.strip() for x in m2.group(4).split(",")]:
            arr = re.match(r"(\w+)\[(\d+)\]", nm)
            fields.append((arr.group(1), f"{typ}[{arr.group(2)}]") if arr else (nm, typ))
    return fields


def _jl_struct(name):
    m = re.search(r"struct " + name + r"\n(.*?)\nend", JL, flags=re.S)
    return [(f.split("::")[0].strip(), f.split("::")[1].strip()) for line in m.group(1).splitlines()
            for f in line.split(";") if "::" in f]


def test_struct_mirrors_match_field_by_field():
    cmap = {"double": "Cdouble", "int64_t": "Int64", "double*": "Ptr{Cdouble}", "qlnlp_model": "QlModel",
            "double[15]": "NTuple{15,Cdouble}"}
    for cname, jname in (("qlnlp_model", "QlModel"), ("qlnlp_problem_desc", "QlProblemDesc"), ("qlnlp_batch_io", "QlBatchIO")):
        c, j = _c_struct(cname), _jl_struct(jname)
        assert [n for n, _ in c] == [n for n, _ in j], (cname, c, j)
        for (n, ct), (_, jt) in zip(c, j):
            assert cmap[ct] == jt, (cname, n, ct, jt)


def test_install_overrides_exactly_the_reference_methods():
    """install! re-defines the seven methods of src/moi.jl:1-33 for ::HybridNLP (so solve() of moi.jl:46-103 needs no
    edit) and makes no false claim about a wrapper type being accepted by solve()."""
    for meth in ("eval_objective", "eval_objective_gradient", "eval_constraint", "eval_constraint_jacobian",
                 "features_available", "initialize", "jacobian_structure"):
        assert re.search(r"\$MOI\." + meth + r"\(prob::HybridNLP", JL), meth
    assert "CudaHybridNLP" not in JL
