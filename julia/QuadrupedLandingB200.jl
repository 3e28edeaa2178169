# QuadrupedLandingB200.jl -- the reference-side binding of libqlnlp.so (include/qlnlp.h).
#
# Drop this file next to the reference's src/moi.jl and `include` it after src/nlp.jl: it defines
# `CudaHybridNLP <: MOI.AbstractNLPEvaluator` with the same seven MOI methods as src/moi.jl:1-33, each a
# one-line `ccall` into the C ABI, so `solve(Z0, CudaHybridNLP(nlp))` runs the unchanged Ipopt/MOI loop of
# src/moi.jl:46-103 against the GPU evaluator.
#
# NOT EXECUTED in the build environment (no julia binary in the image).  The executable stand-in with the
# same semantics is the ctypes wrapper quadruped_landing_b200/evaluator.py, which the test-suite drives.
module QuadrupedLandingB200

using MathOptInterface
const MOI = MathOptInterface

const LIBQLNLP = get(ENV, "LIBQLNLP", "libqlnlp.so")

# mirrors `qlnlp_model` / `qlnlp_problem_desc` of include/qlnlp.h field for field
struct QlModel
    g::Cdouble; mb::Cdouble; mf::Cdouble; lb::Cdouble; l1::Cdouble; l2::Cdouble
end
struct QlProblemDesc
    N::Int64; k_trans::Int64; init_mode::Int64
    model::QlModel
    x0::NTuple{15,Cdouble}; xf::NTuple{15,Cdouble}
    Q::Ptr{Cdouble}; R::Ptr{Cdouble}; q::Ptr{Cdouble}; r::Ptr{Cdouble}; c::Ptr{Cdouble}
end

const JAC_SPARSE_BLOCK = Cint(0)
const JAC_DENSE = Cint(1)
const JAC_SPARSE_TRUE = Cint(2)

check(rc::Cint) = rc == 0 || error("qlnlp error $rc: " * unsafe_string(ccall((:qlnlp_last_error, LIBQLNLP), Cstring, ())))

mutable struct CudaHybridNLP <: MOI.AbstractNLPEvaluator
    handle::Ptr{Cvoid}
    n_nlp::Int
    m_nlp::Int
    nnz::Int
    keep::Vector{Any}     # cost tables referenced by the descriptor during qlnlp_create
end

"""
    CudaHybridNLP(nlp; sparse=true, pattern=:block, device=0)

Build the GPU evaluator from the reference's `HybridNLP` (src/nlp.jl:13-84).  `sparse=false` reports the
dense m_nlp x n_nlp structure exactly like src/moi.jl:31-33; `sparse=true` reports SPARSE_BLOCK, the
column-major filter of the entries `jac_c!` assigns (32,161 instead of 1,327,995 pairs at the default
instance -- this is what removes the 575 s Ipopt spends on the dense structure, src/main.ipynb:724), or with
`pattern=:true` only the 4,840 structurally non-zero entries.
"""
function CudaHybridNLP(nlp; sparse::Bool=true, pattern::Symbol=:block, device::Integer=0)
    N = nlp.N
    Q = Matrix{Cdouble}(undef, 15, N); R = Matrix{Cdouble}(undef, 5, N)
    q = Matrix{Cdouble}(undef, 15, N); r = Matrix{Cdouble}(undef, 5, N); c = Vector{Cdouble}(undef, N)
    for k = 1:N                      # knot-major rows == column-major 15xN / 5xN in Julia
        o = nlp.obj[k]
        Q[:, k] .= o.Q.diag; R[:, k] .= o.R.diag; q[:, k] .= o.q; r[:, k] .= o.r; c[k] = o.c
    end
    m = nlp.model
    desc = Ref(QlProblemDesc(N, nlp.k_trans, nlp.init_mode, QlModel(m.g, m.mb, m.mf, m.lb, m.l1, m.l2),
                             Tuple(nlp.x0), Tuple(nlp.xf), pointer(Q), pointer(R), pointer(q), pointer(r), pointer(c)))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve Q R q r c check(ccall((:qlnlp_create, LIBQLNLP), Cint,
        (Ref{QlProblemDesc}, Cint, Cint, Ref{Ptr{Cvoid}}), desc, device,
        sparse ? (pattern == :true ? JAC_SPARSE_TRUE : JAC_SPARSE_BLOCK) : JAC_DENSE, h))
    n = Ref{Int64}(0); mm = Ref{Int64}(0); nnz = Ref{Int64}(0); nb = Ref{Int64}(0)
    check(ccall((:qlnlp_dims, LIBQLNLP), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}), h[], n, mm, nnz, nb))
    ev = CudaHybridNLP(h[], n[], mm[], nnz[], Any[Q, R, q, r, c])
    finalizer(e -> ccall((:qlnlp_destroy, LIBQLNLP), Cint, (Ptr{Cvoid},), e.handle), ev)
    return ev
end

num_primals(p::CudaHybridNLP) = p.n_nlp          # src/nlp.jl:86
num_duals(p::CudaHybridNLP) = p.m_nlp            # src/nlp.jl:87

# ---- the seven MOI methods of src/moi.jl:1-33 ------------------------------------------------
function MOI.eval_objective(prob::CudaHybridNLP, x)                       # moi.jl:1-3
    f = Ref{Cdouble}(0.0)
    check(ccall((:qlnlp_eval_objective, LIBQLNLP), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ref{Cdouble}), prob.handle, x, f))
    return f[]
end

function MOI.eval_objective_gradient(prob::CudaHybridNLP, grad_f, x)      # moi.jl:5-8
    check(ccall((:qlnlp_eval_objective_gradient, LIBQLNLP), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), prob.handle, x, grad_f))
    return nothing
end

function MOI.eval_constraint(prob::CudaHybridNLP, g, x)                   # moi.jl:10-13
    check(ccall((:qlnlp_eval_constraint, LIBQLNLP), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), prob.handle, x, g))
    return nothing
end

function MOI.eval_constraint_jacobian(prob::CudaHybridNLP, vec, x)        # moi.jl:15-24
    check(ccall((:qlnlp_eval_constraint_jacobian, LIBQLNLP), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), prob.handle, x, vec))
    return nothing
end

MOI.features_available(prob::CudaHybridNLP) = [:Grad, :Jac]               # moi.jl:26-28
MOI.initialize(prob::CudaHybridNLP, features) = nothing                   # moi.jl:30

function MOI.jacobian_structure(prob::CudaHybridNLP)                      # moi.jl:31-33
    rows = Vector{Int64}(undef, prob.nnz); cols = Vector{Int64}(undef, prob.nnz)
    check(ccall((:qlnlp_jacobian_structure, LIBQLNLP), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), prob.handle, rows, cols))
    return collect(zip(rows, cols))                                       # 1-based (row, col), value order
end

# constraint bounds for MOI.NLPBoundsPair.(c_l, c_u) in solve(), src/moi.jl:69-74 / src/nlp.jl:66-69
function constraint_bounds(prob::CudaHybridNLP)
    lb = Vector{Cdouble}(undef, prob.m_nlp); ub = similar(lb)
    check(ccall((:qlnlp_constraint_bounds, LIBQLNLP), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), prob.handle, lb, ub))
    return lb, ub
end

end # module
