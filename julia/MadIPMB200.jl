# MadIPMB200.jl -- Julia side of the B200 backend: MadIPM's KKT system, linear solver and array kernels re-pointed at
# libmadipm_b200.so (C ABI: include/madipm_b200.h) through plain `ccall`.
#
# STATUS: UNTESTED IN JULIA.  There is no Julia toolchain (nor MadNLP / MadIPM sources) in the build or test environment.
# What is checked mechanically (tests/test_julia_shim.py): every `ccall` in this file names an exported function of the
# header with the right number of arguments and matching C types, every VEC_* / status constant equals the header's enum,
# and every MadNLP / MadIPM method of the reference's own KKT system (src/KKT/normalkkt.jl:29-220) and of its linear-solver
# contract (src/utils.jl:54-62,126; src/structure.jl:82; src/solver.jl:360) has a definition here.  The Python twin of the
# same calls (madqp.jl_b200/_lib.py + mpc.py) is what the GPU test-suite runs.
#
# Design.  All device memory belongs to the handle (header, "Ownership").  A Julia vector of the solver is therefore a
# *name*: `B200Vector` = (handle, vector tag, offset, length).  MadNLP's `PrimalVector` / `UnreducedKKTVector` are built
# over such names (same layout as the reference: [x | y | zl | zu] contiguous, structure.jl:130-151), so the control code
# of src/solver.jl that touches vectors directly keeps working on them:
#
#   reference lines                                   | runs                | through
#   --------------------------------------------------+---------------------+--------------------------------------------
#   solver.jl:254-345  mpc!                           | UNCHANGED           | the methods below
#   solver.jl:180-251  affine / Mehrotra / Gondzio    | UNCHANGED           | copyto!(::B200Vector, ::B200Vector), kernels
#   solver.jl:347-403  solve!                         | UNCHANGED           | introduce, update!(stats, solver)
#   linear_solver.jl:6-17 factorize_regularized_system! | UNCHANGED         | set_aug_diagonal_reg!, factorize_wrapper!, is_factorized
#   linear_solver.jl:19-45 solve_system!              | UNCHANGED           | copyto!, solve!, mul!, norm(., Inf)
#   solver.jl:332-333  axpy!(alpha, primal(d), primal(x)) | UNCHANGED       | LinearAlgebra.axpy! -> madipm_vec_axpy
#   solver.jl:334-335  zl_r .+= alpha_d .* dual_lb(d) | UNCHANGED           | broadcast on an index view -> madipm_vec_axpy_at
#   solver.jl:342      adjust_boundary!               | re-pointed (method) | madipm_adjust_boundary
#   solver.jl:338-340  eval_f / eval_cons / eval_grad_f wrappers | re-pointed (methods) | madipm_eval_model (computes Qx once)
#   solver.jl:259-272  jtprod!, get_inf_pr, get_inf_du, get_optimality_gap | re-pointed | madipm_spmv, madipm_phase_residuals
#   solver.jl:6-125    init_starting_point!           | steps 0-2 UNCHANGED, steps 3-end re-pointed | madipm_init_multipliers_and_shift
#   solver.jl:127-178  initialize!                    | re-pointed (method)  | madipm_load_problem (+ the unchanged tail)
#   kernels.jl:1-446   rhs builders, corrections, ratio tests, complementarity, step rules | re-pointed (methods specialised on
#                                                       the solver type, as the reference does for ScaledSparseKKTSystem at kernels.jl:149)
#   structure.jl:76-171 MPCSolver(nlp; ...)           | mirrored below      | vectors are handle names, not allocations
#
# Multi-GPU: none here.  A batch of independent QPs is sharded by the Python harness (madqp.jl_b200/batch.py); a Julia
# caller would create one handle per GPU (cfg.device) from one task each.
module MadIPMB200

using LinearAlgebra
using SparseArrays
import MadNLP
import MadIPM
import NLPModels
import QuadraticModels

const LIB = get(ENV, "MADIPM_B200_LIB", joinpath(@__DIR__, "..", "madqp.jl_b200", "libmadipm_b200.so"))

# ---- constants of include/madipm_b200.h (checked against the header by tests/test_julia_shim.py) -------------------
const MADIPM_OK = Int32(0)
const MADIPM_FACTOR_BREAKDOWN = Int32(1)
const MADIPM_NONFINITE = Int32(2)
const MADIPM_CUDA_ERROR = Int32(3)
const MADIPM_BAD_ARGUMENT = Int32(4)
const MADIPM_INTERNAL = Int32(5)

const MADIPM_KKT_K2 = Int32(0)
const MADIPM_KKT_NORMAL = Int32(1)
const MADIPM_KKT_K2_SCALED = Int32(2)

const MADIPM_OP_A = Int32(0)
const MADIPM_OP_AT = Int32(1)
const MADIPM_OP_Q = Int32(2)

const MADIPM_VEC_X = Int32(0)
const MADIPM_VEC_XL = Int32(1)
const MADIPM_VEC_XU = Int32(2)
const MADIPM_VEC_ZL = Int32(3)
const MADIPM_VEC_ZU = Int32(4)
const MADIPM_VEC_F = Int32(5)
const MADIPM_VEC_JACL = Int32(6)
const MADIPM_VEC_COST = Int32(7)
const MADIPM_VEC_Y = Int32(8)
const MADIPM_VEC_C = Int32(9)
const MADIPM_VEC_RHS = Int32(10)
const MADIPM_VEC_D = Int32(11)
const MADIPM_VEC_P = Int32(12)
const MADIPM_VEC_W1 = Int32(13)
const MADIPM_VEC_W2 = Int32(14)
const MADIPM_VEC_CORR_LB = Int32(15)
const MADIPM_VEC_CORR_UB = Int32(16)
const MADIPM_VEC_REG = Int32(17)
const MADIPM_VEC_PR_DIAG = Int32(18)
const MADIPM_VEC_DU_DIAG = Int32(19)
const MADIPM_VEC_L_DIAG = Int32(20)
const MADIPM_VEC_L_LOWER = Int32(21)
const MADIPM_VEC_U_DIAG = Int32(22)
const MADIPM_VEC_U_LOWER = Int32(23)
const MADIPM_VEC_KKT_NZVAL = Int32(24)
const MADIPM_VEC_CON_SCALE = Int32(25)
const MADIPM_VEC_SCALING = Int32(26)

# ---- handle and errors ------------------------------------------------------------------------------------------------
mutable struct Handle
    ptr::Ptr{Cvoid}
end

last_error(h::Handle) = unsafe_string(ccall((:madipm_last_error, LIB), Cstring, (Ptr{Cvoid},), h.ptr))

"Status codes -> the exceptions MadIPM already handles (SURVEY 8b): 1 is not an error here (is_factorized reports it),
2 -> SolveException (linear_solver.jl:41-43), everything else -> LinearSolverException (ERROR_IN_STEP_COMPUTATION, solver.jl:381-383)."
function check(h::Handle, rc::Int32)
    (rc == MADIPM_OK || rc == MADIPM_FACTOR_BREAKDOWN) && return rc
    rc == MADIPM_NONFINITE && throw(MadNLP.SolveException())
    throw(MadNLP.LinearSolverException("madipm_b200 status $rc: $(last_error(h))"))
end

struct B200Config                       # mirrors madipm_config field for field
    kkt_kind::Int32
    ordering::Int32
    max_width::Int32
    relax_small::Int32
    relax_zeros::NTuple{3,Float64}
    block_panel::Int64
    device::Int32
    batch::Int32
    nd_leaf::Int32
    reserved::Int32
end

struct B200LoadOptions                  # mirrors madipm_load_options
    tol::Float64
    bound_push::Float64
    bound_fac::Float64
    max_gradient::Float64
    scaling::Int32
    reserved::Int32
end

function create_handle(kkt_kind::Int32; device::Integer = 0)
    cfg = Ref{B200Config}()
    ccall((:madipm_default_config, LIB), Cvoid, (Ptr{B200Config},), cfg)
    c = cfg[]
    cfg[] = B200Config(kkt_kind, c.ordering, c.max_width, c.relax_small, c.relax_zeros, c.block_panel, Int32(device), Int32(1),
                       c.nd_leaf, Int32(0))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:madipm_create, LIB), Int32, (Ptr{B200Config}, Ptr{Ptr{Cvoid}}), cfg, out)
    rc == MADIPM_OK || error("madipm_create failed with status $rc")
    h = Handle(out[])
    finalizer(x -> ccall((:madipm_destroy, LIB), Int32, (Ptr{Cvoid},), x.ptr), h)
    return h
end

# ---- the vector type: a name for (a range of) a vector that lives in the handle -------------------------------------
struct B200Vector{T} <: AbstractVector{T}
    h::Handle
    tag::Int32
    off::Int            # 0-based offset inside the handle's vector
    len::Int
end

function B200Vector{T}(h::Handle, tag::Int32) where {T}
    n = Ref{Int64}(0)
    check(h, ccall((:madipm_vec_len, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}), h.ptr, tag, n))
    return B200Vector{T}(h, tag, 0, Int(n[]))
end
B200Vector{T}(::UndefInitializer, n::Integer) where {T} =
    error("B200Vector names device memory owned by a handle; it cannot be allocated (build solvers with MadIPMB200.MPCSolver)")

Base.size(v::B200Vector) = (v.len,)
Base.IndexStyle(::Type{<:B200Vector}) = IndexLinear()
Base.similar(v::B200Vector{T}) where {T} = Vector{T}(undef, v.len)          # host scratch; device vectors are never cloned
Base.similar(v::B200Vector, ::Type{S}, dims::Dims) where {S} = Array{S}(undef, dims)

"Contiguous sub-ranges alias the same device memory (MadNLP builds primal(w), dual(w), dual_lb(w), dual_ub(w) this way)."
MadNLP._madnlp_unsafe_wrap(v::B200Vector{T}, n, shift = 1) where {T} = B200Vector{T}(v.h, v.tag, v.off + shift - 1, n)
Base.view(v::B200Vector{T}, r::UnitRange{Int}) where {T} = B200Vector{T}(v.h, v.tag, v.off + first(r) - 1, length(r))

# host <-> device
function Base.copyto!(dst::Vector{Float64}, src::B200Vector{Float64})
    length(dst) == src.len || throw(DimensionMismatch())
    check(src.h, ccall((:madipm_vec_get_range, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Float64}),
                       src.h.ptr, src.tag, src.off, src.len, dst))
    return dst
end
function Base.copyto!(dst::B200Vector{Float64}, src::Vector{Float64})
    length(src) == dst.len || throw(DimensionMismatch())
    check(dst.h, ccall((:madipm_vec_set_range, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Float64}),
                       dst.h.ptr, dst.tag, dst.off, dst.len, src))
    return dst
end
Base.Array(v::B200Vector{Float64}) = copyto!(Vector{Float64}(undef, v.len), v)
Base.collect(v::B200Vector{Float64}) = Array(v)
# scalar indexing = one device round trip per element: only MadNLP's printing and tests should ever come here
Base.getindex(v::B200Vector{Float64}, i::Int) = Array(B200Vector{Float64}(v.h, v.tag, v.off + i - 1, 1))[1]
function Base.setindex!(v::B200Vector{Float64}, x, i::Int)
    copyto!(B200Vector{Float64}(v.h, v.tag, v.off + i - 1, 1), [Float64(x)])
    return v
end

# device <-> device
function Base.fill!(v::B200Vector{Float64}, x)
    check(v.h, ccall((:madipm_vec_fill_range, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Float64),
                     v.h.ptr, v.tag, v.off, v.len, Float64(x)))
    return v
end
function Base.copyto!(dst::B200Vector{Float64}, src::B200Vector{Float64})
    dst.len == src.len || throw(DimensionMismatch())
    check(dst.h, ccall((:madipm_vec_copy_range, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Int32, Int64, Int64),
                       dst.h.ptr, dst.tag, dst.off, src.tag, src.off, dst.len))
    return dst
end
"axpy!(alpha, x, y): y += alpha x  (solver.jl:28, 332-333)"
function LinearAlgebra.axpy!(alpha::Number, x::B200Vector{Float64}, y::B200Vector{Float64})
    x.len == y.len || throw(DimensionMismatch())
    check(y.h, ccall((:madipm_vec_axpy, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Int32, Int64, Int64, Float64),
                     y.h.ptr, y.tag, y.off, x.tag, x.off, y.len, Float64(alpha)))
    return y
end
"norm(v, Inf) on the device (linear_solver.jl:32-34, solver.jl:174-175, structure.jl:191); other norms through the host."
function LinearAlgebra.norm(v::B200Vector{Float64}, p::Real = 2)
    if p == Inf
        out = Ref{Float64}(0.0)
        check(v.h, ccall((:madipm_vec_norm_inf_range, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Int64, Ptr{Float64}),
                         v.h.ptr, v.tag, v.off, v.len, out))
        return out[]
    end
    return norm(Array(v), p)
end

# Broadcast assignments.  Only the forms src/solver.jl performs on solver vectors are mapped; anything else is an error
# that names the expression instead of silently falling back to scalar indexing.
const BC = Base.Broadcast.Broadcasted
const B200IndexView{T} = SubArray{T,1,B200Vector{T},Tuple{Vector{Int}},false}

Base.Broadcast.materialize!(dest::B200Vector{Float64}, x::Number) = fill!(dest, x)
Base.Broadcast.materialize!(dest::B200Vector{Float64}, x::B200Vector{Float64}) = copyto!(dest, x)
function Base.Broadcast.materialize!(dest::B200Vector{Float64}, bc::BC)
    if bc.f === identity && length(bc.args) == 1
        a = bc.args[1]
        a isa Number && return fill!(dest, a)                      # kkt.reg .= solver.del_w      (solver.jl:16-18)
        a isa B200Vector && return copyto!(dest, a)                # solver.y .= dual(solver.d)   (solver.jl:33)
        a isa Vector{Float64} && return copyto!(dest, a)
    end
    error("MadIPMB200: unsupported broadcast into a device vector: $(bc.f) over $(typeof.(bc.args))")
end

"which index set a view of a device vector selects: 0 = ind_lb, 1 = ind_ub (the only two MadIPM builds, structure.jl:144-151)"
function which_index(v::B200IndexView, kkt)
    idx = parentindices(v)[1]
    idx === kkt.ind_lb && return Int32(0)
    idx === kkt.ind_ub && return Int32(1)
    error("MadIPMB200: index view over an index set that is neither ind_lb nor ind_ub")
end
const INDEX_OWNER = IdDict{Any,Any}()       # index vector => KKT system that registered it (create_kkt_system)

function Base.Broadcast.materialize!(dest::B200IndexView{Float64}, bc::BC)
    # solver.zl_r .+= solver.alpha_d .* MadNLP.dual_lb(solver.d)      (solver.jl:334-335)
    if bc.f === (+) && length(bc.args) == 2 && bc.args[1] === dest && bc.args[2] isa BC
        inner = bc.args[2]
        if inner.f === (*) && length(inner.args) == 2 && inner.args[1] isa Number && inner.args[2] isa B200Vector
            alpha, src = inner.args
            p = parent(dest)
            kkt = INDEX_OWNER[parentindices(dest)[1]]
            check(p.h, ccall((:madipm_vec_axpy_at, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Int64, Float64),
                             p.h.ptr, p.tag, which_index(dest, kkt), src.tag, src.off, Float64(alpha)))
            return dest
        end
    end
    error("MadIPMB200: unsupported broadcast into an index view of a device vector: $(bc.f)")
end

# ---- linear solver: the factorisation lives in the handle ------------------------------------------------------------
"Stands where MadNLP.LDLSolver / CUDSSSolver stand (`linear_solver=` option, utils.jl:72,109).  Constructor contract of
normalkkt.jl:99-101: `LS(csc; opt)`; the matrix argument only carries the pattern the handle already analysed."
mutable struct B200LDLSolver{T} <: MadNLP.AbstractLinearSolver{T}
    h::Union{Nothing,Handle}
    opt::Any
end
struct B200LDLOptions <: MadNLP.AbstractOptions end
B200LDLSolver(csc::SparseMatrixCSC{T}; opt = B200LDLOptions(), kwargs...) where {T} = B200LDLSolver{T}(nothing, opt)
MadNLP.default_options(::Type{B200LDLSolver}) = B200LDLOptions()                               # utils.jl:126, normalkkt.jl:34
MadNLP.is_supported(::Type{B200LDLSolver}, ::Type{Float64}) = true                              # structure.jl:82
MadNLP.is_supported(::Type{B200LDLSolver}, ::Type{T}) where {T} = false
MadNLP.input_type(::Type{B200LDLSolver}) = :csc
MadNLP.introduce(::B200LDLSolver) = "madipm_b200 (supernodal LDL' on one B200, $(unsafe_string(ccall((:madipm_version, LIB), Cstring, ()))))"   # solver.jl:360
MadNLP.is_inertia(::B200LDLSolver) = true
function MadNLP.inertia(ls::B200LDLSolver)
    p, z, n = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    check(ls.h, ccall((:madipm_inertia, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), ls.h.ptr, Int32(0), p, z, n))
    return (Int(p[]), Int(z[]), Int(n[]))
end
"MadNLP.factorize!(linear_solver) (called by factorize_wrapper! after build_kkt!)"
function MadNLP.factorize!(ls::B200LDLSolver)
    nfail = Ref{Int32}(0)
    check(ls.h, ccall((:madipm_factorize, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), ls.h.ptr, nfail))
    return ls
end
"utils.jl:54-62: feeds the retry of linear_solver.jl:8-16"
function MadIPM.is_factorized(ls::B200LDLSolver)
    f = Ref{Int32}(0)
    check(ls.h, ccall((:madipm_is_factorized, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}), ls.h.ptr, f))
    return f[] == 1
end
"MadNLP.solve!(linear_solver, rhs) (normalkkt.jl:196): rhs is a device vector of the KKT order; solved in place"
function MadNLP.solve!(ls::B200LDLSolver, rhs::B200Vector{Float64})
    ptr = Ref{Ptr{Cvoid}}(C_NULL)
    check(ls.h, ccall((:madipm_vec_device_ptr, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Ptr{Cvoid}}), ls.h.ptr, rhs.tag, ptr))
    n = Ref{Int64}(0)
    check(ls.h, ccall((:madipm_vec_len, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}), ls.h.ptr, rhs.tag, n))
    check(ls.h, ccall((:madipm_linear_solve, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64),
                      ls.h.ptr, Ptr{Float64}(ptr[]) + 8 * rhs.off, n[]))
    return rhs
end

# ---- KKT system (method list of src/KKT/normalkkt.jl:1-220) -------------------------------------------------------------
"Fields the control code touches directly (SURVEY 8a T1): reg, pr_diag, du_diag, l_diag, u_diag, l_lower, u_lower,
ind_lb, ind_ub, linear_solver (solver.jl:16-18,360; kernels.jl:135-144; linear_solver.jl:11)."
struct B200KKTSystem{T,LS} <: MadNLP.AbstractKKTSystem{T,B200Vector{T},SparseMatrixCSC{T,Int32},MadNLP.ExactHessian{T,B200Vector{T}}}
    h::Handle
    kind::Int32
    aug_com::SparseMatrixCSC{T,Int32}      # pattern of the assembled KKT matrix (values stay on the device)
    jac::Vector{T}                         # host staging of the Jacobian coordinate values (get_jacobian)
    hess::Vector{T}                        # ... and of the Hessian's (get_hessian)
    reg::B200Vector{T}
    pr_diag::B200Vector{T}
    du_diag::B200Vector{T}
    l_diag::B200Vector{T}
    u_diag::B200Vector{T}
    l_lower::B200Vector{T}
    u_lower::B200Vector{T}
    linear_solver::LS
    ind_ineq::Vector{Int}
    ind_lb::Vector{Int}
    ind_ub::Vector{Int}
    n::Int
    m::Int
end
# the three flavours the `kkt_system=` option selects (utils.jl:71,108; runtests.jl:102-115,139-180)
struct B200SparseKKTSystem end           # MadNLP.SparseKKTSystem        (K2)
struct B200NormalKKTSystem end           # MadIPM.NormalKKTSystem        (LP, normal equations)
struct B200ScaledSparseKKTSystem end     # MadNLP.ScaledSparseKKTSystem  (K2.5)
kkt_kind(::Type{B200SparseKKTSystem}) = MADIPM_KKT_K2
kkt_kind(::Type{B200NormalKKTSystem}) = MADIPM_KKT_NORMAL
kkt_kind(::Type{B200ScaledSparseKKTSystem}) = MADIPM_KKT_K2_SCALED

function symbolic(h::Handle, which::Integer)
    len = Ref{Int64}(0)
    check(h, ccall((:madipm_get_symbolic, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}, Int64, Ptr{Int64}), h.ptr, Int32(which), C_NULL, 0, len))
    out = Vector{Int64}(undef, len[])
    check(h, ccall((:madipm_get_symbolic, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}, Int64, Ptr{Int64}), h.ptr, Int32(which), out, len[], len))
    return out
end
const SYM_KKT_COLPTR = 9
const SYM_KKT_ROWVAL = 10

"Signature of src/KKT/normalkkt.jl:29-37."
function MadNLP.create_kkt_system(
    ::Type{KT},
    cb::MadNLP.AbstractCallback{T},
    ind_cons,
    linear_solver::Type;
    opt_linear_solver = MadNLP.default_options(linear_solver),
    hessian_approximation = MadNLP.ExactHessian,
    qn_options = MadNLP.QuasiNewtonOptions(),
    device::Integer = 0,
) where {T,KT<:Union{B200SparseKKTSystem,B200NormalKKTSystem,B200ScaledSparseKKTSystem}}
    kind = kkt_kind(KT)
    n, m = cb.nvar, cb.ncon
    ind_ineq = Vector{Int}(ind_cons.ind_ineq)
    n_slack = length(ind_ineq)
    if kind == MADIPM_KKT_NORMAL && cb.nnzh > 0
        error("The KKT system NormalKKTSystem supports only linear programs.")                 # normalkkt.jl:45-48
    end
    h = create_handle(kind; device = device)
    jI = Vector{Int32}(undef, cb.nnzj); jJ = Vector{Int32}(undef, cb.nnzj)
    MadNLP._jac_sparsity_wrapper!(cb, jI, jJ)                                                  # normalkkt.jl:51-53
    ineq64 = Vector{Int64}(ind_ineq)
    if kind == MADIPM_KKT_NORMAL
        check(h, ccall((:madipm_set_pattern_normal, LIB), Int32,
                       (Ptr{Cvoid}, Int32, Int32, Int64, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int64}),
                       h.ptr, Int32(n), Int32(m), Int64(length(jI)), jI, jJ, Int32(n_slack), ineq64))
    else
        hI = Vector{Int32}(undef, cb.nnzh); hJ = Vector{Int32}(undef, cb.nnzh)
        MadNLP._hess_sparsity_wrapper!(cb, hI, hJ)
        check(h, ccall((:madipm_set_pattern_k2, LIB), Int32,
                       (Ptr{Cvoid}, Int32, Int32, Int64, Ptr{Int32}, Ptr{Int32}, Int64, Ptr{Int32}, Ptr{Int32}, Int32, Ptr{Int64}),
                       h.ptr, Int32(n), Int32(m), Int64(length(hI)), hI, hJ, Int64(length(jI)), jI, jJ, Int32(n_slack), ineq64))
    end
    lb = Vector{Int64}(ind_cons.ind_lb); ub = Vector{Int64}(ind_cons.ind_ub)                   # structure.jl:144-151
    check(h, ccall((:madipm_set_bounds_index, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int64}, Int32, Ptr{Int64}),
                   h.ptr, Int32(length(lb)), lb, Int32(length(ub)), ub))
    check(h, ccall((:madipm_analyze, LIB), Int32, (Ptr{Cvoid},), h.ptr))                       # ordering + symbolic LDL': normalkkt.jl:99-101
    # pattern of aug_com, lower triangle, for code that inspects it (values never leave the device unless asked)
    colptr = Vector{Int32}(symbolic(h, SYM_KKT_COLPTR) .+ 1)
    rowval = Vector{Int32}(symbolic(h, SYM_KKT_ROWVAL) .+ 1)
    N = length(colptr) - 1
    aug_com = SparseMatrixCSC{T,Int32}(N, N, colptr, rowval, zeros(T, length(rowval)))
    ls = linear_solver(aug_com; opt = opt_linear_solver)
    ls.h = h
    V = B200Vector{T}
    kkt = B200KKTSystem{T,typeof(ls)}(
        h, kind, aug_com, zeros(T, cb.nnzj), zeros(T, cb.nnzh),
        V(h, MADIPM_VEC_REG), V(h, MADIPM_VEC_PR_DIAG), V(h, MADIPM_VEC_DU_DIAG), V(h, MADIPM_VEC_L_DIAG), V(h, MADIPM_VEC_U_DIAG),
        V(h, MADIPM_VEC_L_LOWER), V(h, MADIPM_VEC_U_LOWER), ls,
        ind_ineq, ind_cons.ind_lb, ind_cons.ind_ub, n + n_slack, m)
    INDEX_OWNER[ind_cons.ind_lb] = kkt
    INDEX_OWNER[ind_cons.ind_ub] = kkt
    return kkt
end

MadNLP.num_variables(kkt::B200KKTSystem) = kkt.n                                               # normalkkt.jl:128
MadNLP.get_jacobian(kkt::B200KKTSystem) = kkt.jac                                              # normalkkt.jl:129
MadNLP.get_hessian(kkt::B200KKTSystem) = kkt.hess                                              # normalkkt.jl:130
function MadNLP.is_inertia_correct(kkt::B200KKTSystem, num_pos, num_zero, num_neg)             # normalkkt.jl:132-134
    kkt.kind == MADIPM_KKT_NORMAL && return (num_zero == 0) && (num_pos == kkt.m)
    return (num_zero == 0) && (num_pos == kkt.n) && (num_neg == kkt.m)
end
"normalkkt.jl:136-147"
MadNLP.initialize!(kkt::B200KKTSystem) = (check(kkt.h, ccall((:madipm_kkt_initialize, LIB), Int32, (Ptr{Cvoid},), kkt.h.ptr)); nothing)
"normalkkt.jl:149-160 (+ compress_hessian!): the coordinate values the callbacks wrote into get_jacobian / get_hessian go to the
device, slack columns (-1) and the CSR / CSC compressions are produced there (one call serves both)."
function MadNLP.compress_jacobian!(kkt::B200KKTSystem)
    check(kkt.h, ccall((:madipm_set_values, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), kkt.h.ptr, kkt.hess, kkt.jac))
    return
end
MadNLP.compress_hessian!(kkt::B200KKTSystem) = MadNLP.compress_jacobian!(kkt)
"normalkkt.jl:162-164: y = A' x"
function MadNLP.jtprod!(y::B200Vector{Float64}, kkt::B200KKTSystem, x::B200Vector{Float64})
    check(kkt.h, ccall((:madipm_spmv, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), kkt.h.ptr, MADIPM_OP_AT, y.tag, x.tag))
    return y
end
"normalkkt.jl:166-180 (normal equations) / MadNLP's K2 and K2.5 build_kkt!"
MadNLP.build_kkt!(kkt::B200KKTSystem) = (check(kkt.h, ccall((:madipm_build_kkt, LIB), Int32, (Ptr{Cvoid},), kkt.h.ptr)); nothing)
"normalkkt.jl:182-205 and the K2 / K2.5 solve! of MadNLP: reduce_rhs!, solve, finish_aug_solve! in place on w"
function MadNLP.solve!(kkt::B200KKTSystem, w::MadNLP.AbstractKKTVector)
    check(kkt.h, ccall((:madipm_kkt_solve, LIB), Int32, (Ptr{Cvoid}, Int32), kkt.h.ptr, MadNLP.full(w).tag))
    return w
end
"normalkkt.jl:207-219"
function LinearAlgebra.mul!(w::MadNLP.AbstractKKTVector{T}, kkt::B200KKTSystem, v::MadNLP.AbstractKKTVector, alpha = one(T), beta = zero(T)) where {T}
    check(kkt.h, ccall((:madipm_kkt_mul, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Float64, Float64),
                       kkt.h.ptr, MadNLP.full(w).tag, MadNLP.full(v).tag, Float64(alpha), Float64(beta)))
    return w
end
"kernels.jl:128-165 (both the plain and the ScaledSparseKKTSystem method: the handle knows its kind)"
function MadIPM.set_aug_diagonal_reg!(kkt::B200KKTSystem{T}, solver::MadNLP.AbstractMadNLPSolver{T}) where {T}
    check(kkt.h, ccall((:madipm_set_aug_diagonal_reg, LIB), Int32, (Ptr{Cvoid}, Float64, Float64), kkt.h.ptr, solver.del_w, solver.del_c))
    return
end

# ---- model and callback stand-ins (the reference evaluates the model on the device too: scripts/qp_gpu.jl:29-40) -----
"The QuadraticModel as MPCSolver sees it: vector type = B200Vector (structure.jl:5), data on the host for the index
computations of structure.jl:94-101, evaluation on the device."
struct B200QPModel{T} <: NLPModels.AbstractNLPModel{T,B200Vector{T}}
    meta::NLPModels.NLPModelMeta{T,Vector{T}}
    counters::NLPModels.Counters
    qp::QuadraticModels.QuadraticModel
end
"What MadIPM reads from `solver.cb` (structure.jl:175 obj_scale; sparsity wrappers in create_kkt_system)."
struct B200Callback{T} <: MadNLP.AbstractCallback{T}
    qp::QuadraticModels.QuadraticModel
    nvar::Int
    ncon::Int
    nnzj::Int
    nnzh::Int
    obj_scale::Base.RefValue{T}
    con_scale::Vector{T}
end
function MadNLP._jac_sparsity_wrapper!(cb::B200Callback, I::AbstractVector, J::AbstractVector)
    NLPModels.jac_structure!(cb.qp, I, J)
    return
end
function MadNLP._hess_sparsity_wrapper!(cb::B200Callback, I::AbstractVector, J::AbstractVector)
    NLPModels.hess_structure!(cb.qp, I, J)
    return
end

const B200Solver{T} = MadIPM.MPCSolver{T,B200Vector{T}}
handle(solver::B200Solver) = solver.kkt.h

"structure.jl:76-171 with vectors that are names of the handle's vectors instead of allocations.  `kkt_system` is one of
B200SparseKKTSystem (default), B200NormalKKTSystem, B200ScaledSparseKKTSystem; other options as MadIPM.MPCSolver."
function MPCSolver(qp::QuadraticModels.QuadraticModel{T}; kkt_system = B200SparseKKTSystem, device::Integer = 0, kwargs...) where {T}
    nlp = B200QPModel{T}(qp.meta, NLPModels.Counters(), qp)
    options = MadIPM.load_options(qp; kkt_system = kkt_system, linear_solver = B200LDLSolver, kwargs...)   # utils.jl:99-147
    ipm_opt = options.interior_point
    @assert MadNLP.is_supported(ipm_opt.linear_solver, T)                                      # structure.jl:82
    cnt = MadNLP.MadNLPCounters(start_time = time())
    cb = B200Callback{T}(qp, NLPModels.get_nvar(qp), NLPModels.get_ncon(qp), NLPModels.get_nnzj(qp), NLPModels.get_nnzh(qp),
                         Ref(one(T)), ones(T, NLPModels.get_ncon(qp)))
    ind_cons = MadNLP.get_index_constraints(                                                    # structure.jl:94-101
        NLPModels.get_lvar(qp), NLPModels.get_uvar(qp), NLPModels.get_lcon(qp), NLPModels.get_ucon(qp);
        fixed_variable_treatment = ipm_opt.fixed_variable_treatment, equality_treatment = ipm_opt.equality_treatment)
    ind_lb, ind_ub = ind_cons.ind_lb, ind_cons.ind_ub
    ns = length(ind_cons.ind_ineq); nx = NLPModels.get_nvar(qp); n = nx + ns; m = NLPModels.get_ncon(qp)
    nlb, nub = length(ind_lb), length(ind_ub)
    kkt = MadNLP.create_kkt_system(ipm_opt.kkt_system, cb, ind_cons, ipm_opt.linear_solver;
                                   opt_linear_solver = options.linear_solver, device = device)  # structure.jl:115-121
    h = kkt.h
    V = B200Vector{T}; VI = Vector{Int}
    primal_vec(tag) = begin                                                                     # MadNLP.PrimalVector(VT, nx, ns, ind_lb, ind_ub)
        values = V(h, tag)
        MadNLP.PrimalVector{T,V,VI}(values, view(values, ind_lb), view(values, ind_ub),
                                    MadNLP._madnlp_unsafe_wrap(values, nx), MadNLP._madnlp_unsafe_wrap(values, ns, nx + 1))
    end
    kkt_vec(tag) = begin                                                                        # MadNLP.UnreducedKKTVector(VT, n, m, nlb, nub, ind_lb, ind_ub)
        values = V(h, tag)
        xp = MadNLP._madnlp_unsafe_wrap(values, n)
        MadNLP.UnreducedKKTVector(values, MadNLP._madnlp_unsafe_wrap(values, n + m), xp, view(xp, ind_lb), view(xp, ind_ub),
                                  MadNLP._madnlp_unsafe_wrap(values, m, n + 1), MadNLP._madnlp_unsafe_wrap(values, nlb, n + m + 1),
                                  MadNLP._madnlp_unsafe_wrap(values, nub, n + m + nlb + 1))
    end
    x, xl, xu = primal_vec(MADIPM_VEC_X), primal_vec(MADIPM_VEC_XL), primal_vec(MADIPM_VEC_XU)   # structure.jl:123-128
    zl, zu, f = primal_vec(MADIPM_VEC_ZL), primal_vec(MADIPM_VEC_ZU), primal_vec(MADIPM_VEC_F)
    d, p = kkt_vec(MADIPM_VEC_D), kkt_vec(MADIPM_VEC_P)                                          # structure.jl:130-133
    _w1, _w2 = kkt_vec(MADIPM_VEC_W1), kkt_vec(MADIPM_VEC_W2)
    correction_lb, correction_ub = V(h, MADIPM_VEC_CORR_LB), V(h, MADIPM_VEC_CORR_UB)            # structure.jl:136-142
    jacl, y, c, rhs = V(h, MADIPM_VEC_JACL), V(h, MADIPM_VEC_Y), V(h, MADIPM_VEC_C), V(h, MADIPM_VEC_RHS)
    full = MadNLP.full
    x_lr = view(full(x), ind_lb); x_ur = view(full(x), ind_ub)                                   # structure.jl:144-151
    xl_r = view(full(xl), ind_lb); xu_r = view(full(xu), ind_ub)
    zl_r = view(full(zl), ind_lb); zu_r = view(full(zu), ind_ub)
    dx_lr = view(d.xp, ind_lb); dx_ur = view(d.xp, ind_ub)
    cnt.init_time = time() - cnt.start_time
    class = iszero(NLPModels.get_nnzh(qp)) ? MadIPM.LinearProgram() : MadIPM.QuadraticProgram()
    return MadIPM.MPCSolver(                                                                    # structure.jl:158-171, same order
        nlp, class, cb, kkt, ipm_opt, cnt, options.logger, n, m, nlb, nub,
        x, y, zl, zu, xl, xu, zero(T), f, c, jacl, d, p, _w1, _w2, correction_lb, correction_ub, rhs,
        ind_cons.ind_ineq, ind_cons.ind_fixed, ind_cons.ind_llb, ind_cons.ind_uub, ind_cons.ind_lb, ind_cons.ind_ub,
        x_lr, x_ur, xl_r, xu_r, zl_r, zu_r, dx_lr, dx_ur,
        zero(T), zero(T), zero(T), zero(T), zero(T), zero(T), zero(T), zero(T), zero(T), zero(T), MadNLP.INITIAL)
end

# ---- solver.jl:127-178 initialize!: problem loading, scaling and first evaluation on the device ---------------------
function MadIPM.initialize!(solver::B200Solver{T}) where {T}
    opt = solver.opt
    qp = solver.cb.qp
    h = handle(solver)
    hess_V = Vector{T}(undef, solver.cb.nnzh); jac_V = Vector{T}(undef, solver.cb.nnzj)
    NLPModels.jac_coord!(qp, qp.meta.x0, jac_V)
    NLPModels.hess_coord!(qp, qp.meta.x0, hess_V)
    lopt = Ref(B200LoadOptions(opt.bound_relax_factor, opt.bound_push, opt.bound_fac, 100.0, opt.scaling ? Int32(1) : Int32(0), Int32(0)))
    obj_scale = Ref{Float64}(1.0); norm_b = Ref{Float64}(0.0)
    # MadNLP.initialize!(cb, x, xl, xu, y, rhs, ind_ineq; ...) + set_scaling! + eval_jac / eval_lag_hess + compress_*  (solver.jl:131-170)
    check(h, ccall((:madipm_load_problem, LIB), Int32,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Ptr{Float64}, Ptr{B200LoadOptions}, Ptr{Float64}, Ptr{Float64}),
                   h.ptr, qp.meta.x0, qp.meta.lvar, qp.meta.uvar, qp.meta.y0, qp.meta.lcon, qp.meta.ucon, qp.data.c, hess_V, jac_V,
                   lopt, obj_scale, norm_b))
    solver.cb.obj_scale[] = obj_scale[]
    copyto!(solver.cb.con_scale, B200Vector{T}(h, MADIPM_VEC_CON_SCALE))
    fill!(solver.jacl, zero(T))                                                                 # solver.jl:145
    MadNLP.initialize!(solver.kkt)                                                              # solver.jl:162
    MadIPM.init_regularization!(solver, solver.opt.regularization)                              # solver.jl:163
    solver.obj_val = MadNLP.eval_f_wrapper(solver, solver.x)                                    # solver.jl:166-170 (one device pass)
    solver.norm_b = norm_b[]                                                                    # solver.jl:173
    solver.norm_c = norm(MadNLP.primal(solver.f), Inf)                                          # solver.jl:174
    MadIPM.init_starting_point!(solver)                                                         # solver.jl:177
    solver.mu = opt.mu_init
    return MadNLP.REGULAR
end

"solver.jl:6-125: steps 0-2 exactly as the reference writes them (they only use methods defined above); steps 3-end (bound
multipliers, shifts, projection, the four asserts) in one device call."
function MadIPM.init_starting_point!(solver::B200Solver{T}) where {T}
    x = MadNLP.primal(solver.x)
    solver.kkt.reg .= solver.del_w
    solver.kkt.pr_diag .= solver.del_w
    solver.kkt.du_diag .= solver.del_c
    MadNLP.factorize_wrapper!(solver)
    MadIPM.set_initial_primal_rhs!(solver)
    MadIPM.solve_system!(solver.d, solver, solver.p)
    axpy!(1.0, MadNLP.primal(solver.d), x)
    MadIPM.set_initial_dual_rhs!(solver)
    MadIPM.solve_system!(solver.d, solver, solver.p)
    solver.y .= MadNLP.dual(solver.d)
    h = handle(solver)
    rc = ccall((:madipm_init_multipliers_and_shift, LIB), Int32, (Ptr{Cvoid}, Float64), h.ptr, solver.opt.bound_fac)
    rc == MADIPM_INTERNAL && throw(AssertionError(last_error(h)))                               # the @asserts of solver.jl:119-122
    check(h, rc)
    return
end

"MadNLP.factorize_wrapper!(solver): build_kkt! + factorize! (linear_solver.jl:10, solver.jl:21)"
function MadNLP.factorize_wrapper!(solver::B200Solver)
    MadNLP.build_kkt!(solver.kkt)
    MadNLP.factorize!(solver.kkt.linear_solver)
    return
end

# ---- solver.jl:338-340 and :166-170: the three callbacks share one device pass (Qx is computed once) ---------------
function MadNLP.eval_f_wrapper(solver::B200Solver{T}, x) where {T}
    obj = Ref{Float64}(0.0)
    h = handle(solver)
    check(h, ccall((:madipm_eval_model, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}), h.ptr, Float64(solver.cb.qp.data.c0), obj))
    return obj[]                                 # scaled objective, as MadNLP.eval_f_wrapper returns it; f and c are up to date as well
end
MadNLP.eval_cons_wrapper!(solver::B200Solver, c, x) = c          # written by eval_f_wrapper's device pass
MadNLP.eval_grad_f_wrapper!(solver::B200Solver, f, x) = f        # ditto
MadNLP.eval_jac_wrapper!(solver::B200Solver, kkt, x) = nothing   # constant for LP/QP: values were loaded by initialize!
MadNLP.eval_lag_hess_wrapper!(solver::B200Solver, kkt, x, y; kwargs...) = nothing
"solver.jl:342"
function MadNLP.adjust_boundary!(x_lr::B200IndexView{T}, xl_r, x_ur, xu_r, mu) where {T}
    p = parent(x_lr)
    check(p.h, ccall((:madipm_adjust_boundary, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), p.h.ptr, Ref{Float64}(mu)))
    return
end

# ---- solver.jl:259-272: convergence measures --------------------------------------------------------------------------
# mpc! calls jtprod!(jacl, kkt, y), get_inf_pr(c), get_inf_du(f, zl, zu, jacl, 1.0) and get_optimality_gap(solver) one after
# the other; the device computes all of them in one pass (madipm_phase_residuals).  Unchanged mpc! therefore runs that pass
# three times per iteration (15 us each); a caller that edits mpc! can call `residuals` once.
function residuals(h::Handle)
    out = Vector{Float64}(undef, 4)
    check(h, ccall((:madipm_phase_residuals, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), h.ptr, out))
    return (out[1], out[2], out[3], out[4])       # max|c|, max|f - zl + zu + A'y|, max complementarity product, max|dx|
end
MadNLP.get_inf_pr(c::B200Vector{Float64}) = residuals(c.h)[1]
MadNLP.get_inf_du(f::B200Vector{Float64}, zl, zu, jacl, sd) = residuals(f.h)[2]
MadIPM.get_optimality_gap(solver::B200Solver) = residuals(handle(solver))[3]                    # kernels.jl:435-446

# ---- kernels.jl: one device call each -----------------------------------------------------------------------------------
MadIPM.set_initial_primal_rhs!(solver::B200Solver) =                                            # kernels.jl:1-9
    (check(handle(solver), ccall((:madipm_set_initial_primal_rhs, LIB), Int32, (Ptr{Cvoid},), handle(solver).ptr)); nothing)
MadIPM.set_initial_dual_rhs!(solver::B200Solver) =                                              # kernels.jl:11-19
    (check(handle(solver), ccall((:madipm_set_initial_dual_rhs, LIB), Int32, (Ptr{Cvoid},), handle(solver).ptr)); nothing)
MadIPM.set_predictive_rhs!(solver::B200Solver, kkt::MadNLP.AbstractKKTSystem) =                 # kernels.jl:21-41
    (check(handle(solver), ccall((:madipm_set_predictive_rhs, LIB), Int32, (Ptr{Cvoid},), handle(solver).ptr)); nothing)
function MadIPM.set_correction_rhs!(solver::B200Solver, kkt::MadNLP.AbstractKKTSystem, mu::Float64, correction_lb::AbstractVector{Float64},
                                    correction_ub::AbstractVector{Float64}, ind_lb, ind_ub)     # kernels.jl:43-61
    check(handle(solver), ccall((:madipm_set_correction_rhs, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}), handle(solver).ptr, Ref{Float64}(mu)))
    return
end
function MadIPM.get_correction!(solver::B200Solver, correction_lb, correction_ub)               # kernels.jl:63-75
    check(handle(solver), ccall((:madipm_get_correction, LIB), Int32, (Ptr{Cvoid},), handle(solver).ptr))
    return
end
function MadIPM.set_extra_correction!(solver::B200Solver, correction_lb, correction_ub, alpha_p, alpha_d, beta_min, beta_max, mu)   # kernels.jl:78-126
    check(handle(solver), ccall((:madipm_set_extra_correction, LIB), Int32,
                                (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ptr{Float64}),
                                handle(solver).ptr, Ref{Float64}(alpha_p), Ref{Float64}(alpha_d), Float64(beta_min), Float64(beta_max), Ref{Float64}(mu)))
    return
end
"kernels.jl:171-190: out = (affine complementarity at the given steps, current complementarity)"
function complementarity(solver::B200Solver, alpha_p, alpha_d)
    out = Vector{Float64}(undef, 2)
    check(handle(solver), ccall((:madipm_complementarity, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                                handle(solver).ptr, Ref{Float64}(alpha_p), Ref{Float64}(alpha_d), out))
    return out
end
MadIPM.get_complementarity_measure(solver::B200Solver) = complementarity(solver, 0.0, 0.0)[2]                        # kernels.jl:171-190
MadIPM.get_affine_complementarity_measure(solver::B200Solver, alpha_p, alpha_d) = complementarity(solver, alpha_p, alpha_d)[1]   # kernels.jl:192-224
"kernels.jl:242-305: the four arg-min reductions in one pass; out8 = (alpha_xl, alpha_xu, alpha_zl, alpha_zu, four blocking indices)"
function MadIPM.get_fraction_to_boundary_step(solver::B200Solver, tau)
    out = Vector{Float64}(undef, 8)
    check(handle(solver), ccall((:madipm_ratio_test, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), handle(solver).ptr, Ref{Float64}(tau), out))
    return min(out[1], out[2]), min(out[3], out[4])
end
"kernels.jl:325-374 without scalar reads of device memory"
function MadIPM.update_step!(rule::MadIPM.MehrotraAdaptiveStep, solver::B200Solver)
    out = Vector{Float64}(undef, 2)
    check(handle(solver), ccall((:madipm_mehrotra_step, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}), handle(solver).ptr, Float64(rule.gamma_f), out))
    solver.alpha_p = out[1]
    solver.alpha_d = out[2]
    return
end

# ---- results (solve! ends with MadNLP.update!(stats, solver), solver.jl:399) --------------------------------------------
function MadNLP.update!(stats::MadNLP.MadNLPExecutionStats, solver::B200Solver{T}) where {T}
    nx = solver.cb.nvar
    stats.status = solver.status
    stats.solution = Array(MadNLP.full(solver.x))[1:nx]
    stats.multipliers = Array(solver.y) .* solver.cb.con_scale ./ solver.cb.obj_scale[]
    stats.objective = solver.obj_val / solver.cb.obj_scale[]
    stats.iter = solver.cnt.k
    stats.primal_feas = solver.inf_pr
    stats.dual_feas = solver.inf_du
    stats.counters = solver.cnt
    return stats
end

# ---- preprocessing: scale_qp of scripts/common.jl:57-100 ------------------------------------------------------------------
struct B200ScaleOptions                 # mirrors madipm_scale_options
    device::Int32
    max_iter::Int32
    tol::Float64
    launches::Int64
end
"Ruiz equilibration of the model's arrays IN PLACE (where the reference calls HSL.mc77); returns (Dr, Dc)."
function scale_qp!(nvar, ncon, jac_I::Vector{Int32}, jac_J::Vector{Int32}, jac_V::Vector{Float64},
                   hess_I::Vector{Int32}, hess_J::Vector{Int32}, hess_V::Vector{Float64},
                   c, lvar, uvar, lcon, ucon, x0, y0; max_iter = 10, tol = 0.0, device = 0)
    opt = Ref{B200ScaleOptions}()
    ccall((:madipm_default_scale_options, LIB), Cvoid, (Ptr{B200ScaleOptions},), opt)
    opt[] = B200ScaleOptions(Int32(device), Int32(max_iter), Float64(tol), 0)
    Dr = Vector{Float64}(undef, ncon); Dc = Vector{Float64}(undef, nvar)
    sweeps = Ref{Int32}(0); err = Ref{Float64}(0.0)
    rc = ccall((:madipm_scale_qp, LIB), Int32,
               (Ptr{B200ScaleOptions}, Int32, Int32, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Int32}, Ptr{Float64}),
               opt, Int32(nvar), Int32(ncon), Int64(length(jac_V)), jac_I, jac_J, jac_V, Int64(length(hess_V)), hess_I, hess_J, hess_V,
               c, lvar, uvar, lcon, ucon, x0, y0, Dr, Dc, sweeps, err)
    rc == MADIPM_OK || error("madipm_scale_qp: " * unsafe_string(ccall((:madipm_scale_last_error, LIB), Cstring, ())))
    return Dr, Dc
end

# ---- preprocessing: standard_form_qp of scripts/common.jl:156-288 ---------------------------------------------------------
"""
    standard_form_arrays(nvar, ncon, jac_I, jac_J, jac_V, c, lvar, uvar, lcon, ucon, x0, y0) -> NamedTuple

The arrays of `standard_form_qp(qp)` (slacks for the inequality rows, range bounds moved into equality rows), computed on
the device; the caller wraps them in a `QuadraticModel` as common.jl:262-287 does (the Hessian triplets are unchanged).
"""
function standard_form_arrays(nvar, ncon, jac_I::Vector{Int32}, jac_J::Vector{Int32}, jac_V::Vector{Float64},
                              c::Vector{Float64}, lvar::Vector{Float64}, uvar::Vector{Float64}, lcon::Vector{Float64},
                              ucon::Vector{Float64}, x0::Vector{Float64}, y0::Vector{Float64}; device = 0)
    nnzj = length(jac_V)
    cap_var = 2 * (nvar + ncon); cap_con = nvar + 2 * ncon; cap_nz = nnzj + ncon + 2 * (nvar + ncon)
    oI = Vector{Int32}(undef, cap_nz); oJ = Vector{Int32}(undef, cap_nz); oV = Vector{Float64}(undef, cap_nz)
    oc = Vector{Float64}(undef, cap_var); olv = Vector{Float64}(undef, cap_var); ouv = Vector{Float64}(undef, cap_var)
    ox0 = Vector{Float64}(undef, cap_var)
    olc = Vector{Float64}(undef, cap_con); ouc = Vector{Float64}(undef, cap_con); oy0 = Vector{Float64}(undef, cap_con)
    ns = Ref{Int32}(0); nw = Ref{Int32}(0); launches = Ref{Int64}(0)
    rc = ccall((:madipm_standard_form_qp, LIB), Int32,
               (Int32, Int32, Int32, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64},
                Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int64}),
               Int32(device), Int32(nvar), Int32(ncon), Int64(nnzj), jac_I, jac_J, jac_V, c, lvar, uvar, lcon, ucon, x0, y0, ns, nw,
               oI, oJ, oV, oc, olv, ouv, ox0, olc, ouc, oy0, launches)
    rc == MADIPM_OK || error("madipm_standard_form_qp: " * unsafe_string(ccall((:madipm_scale_last_error, LIB), Cstring, ())))
    nv = nvar + ns[] + nw[]; nc = ncon + nw[]; nz = nnzj + ns[] + 2 * nw[]
    return (nvar = nv, ncon = nc, ns = Int(ns[]), nw = Int(nw[]), jac_I = oI[1:nz], jac_J = oJ[1:nz], jac_V = oV[1:nz],
            c = oc[1:nv], lvar = olv[1:nv], uvar = ouv[1:nv], x0 = ox0[1:nv], lcon = olc[1:nc], ucon = ouc[1:nc], y0 = oy0[1:nc])
end

# ---- scheduling of the small-front tree kernels (no reference counterpart) --------------------------------------------------
"Dataflow schedule (per-task dependency counters, default) or height barriers; returns whether dataflow is in effect."
function set_tree_dataflow!(h::Handle, on::Bool)
    eff = Ref{Int32}(0)
    check(h, ccall((:madipm_set_tree_dataflow, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{Int32}), h.ptr, Int32(on), eff))
    return eff[] != 0
end

# ---- hang guards (see the header) -----------------------------------------------------------------------------------------
function barrier_state(h::Handle)
    out = Vector{Int64}(undef, 8)
    ccall((:madipm_debug_barrier_state, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}), h.ptr, out)
    return (aborted = out[1], kernel = out[2], op = out[3], height = out[4], block = out[5], grid = out[6], seen = out[7], target = out[8])
end
synchronize(h::Handle) = check(h, ccall((:madipm_synchronize, LIB), Int32, (Ptr{Cvoid},), h.ptr))

"""
    solve(qp::QuadraticModel; kwargs...) -> MadNLPExecutionStats

`MadIPM.madipm(qp; ...)` on one B200: builds the solver above and runs the UNCHANGED `MadIPM.solve!` (solver.jl:347-403).
"""
function solve(qp::QuadraticModels.QuadraticModel; kwargs...)
    solver = MPCSolver(qp; kwargs...)
    return MadIPM.solve!(solver)
end

end # module
