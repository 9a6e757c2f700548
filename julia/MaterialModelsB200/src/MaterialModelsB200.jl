# MaterialModelsB200.jl — the binding a MaterialModels.jl maintainer would add: batched methods of the
# EXISTING generics (`material_response`, `initial_material_state`, `get_cache`) that `ccall` libmmb200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build container has no Julia. It is kept literal and thin so that
# it can be checked by reading against include/mmb200.h.  Arrays are CUDA.jl `CuArray`s (device pointers) or
# plain `Array`s (host pointers -> the pipelined mmh_* entry points).  AoS layout is what
# `reinterpret(Float64, ::Vector{SymmetricTensor{2,3,Float64,6}})` gives, so existing assembler arrays pass
# through unchanged; SoA (`(nc, N)` stored component-major, i.e. an `N x nc` Julia matrix) is the fast path.
module MaterialModelsB200

using MaterialModels, Tensors, CUDA
import MaterialModels: material_response, initial_material_state, get_cache

const libmmb200 = get(ENV, "LIBMMB200", "libmmb200.so")
const MM_LAYOUT_SOA, MM_LAYOUT_AOS = Cint(0), Cint(1)

struct MMPlastic;  E::Cdouble; nu::Cdouble; sigma_y::Cdouble; H::Cdouble; r::Cdouble; kappa_inf::Cdouble; alpha_inf::Cdouble; end
struct MMLinearElastic; E::Cdouble; nu::Cdouble; end
struct MMHyper; lambda::Cdouble; mu::Cdouble; c2::Cdouble; c3::Cdouble; end
struct MMXuNeedleman; Phi_n::Cdouble; Phi_t::Cdouble; delta_n::Cdouble; delta_t::Cdouble; Delta_n_star::Cdouble; end
struct MMNewtonOpts; tol::Cdouble; max_iter::Int32; pad::Int32; end
MMPlastic(m::Plastic) = MMPlastic(m.E, m.ν, m.σ_y, m.H, m.r, m.κ_∞, m.α_∞)

check(rc, what) = rc == 0 || error("$what failed: " * unsafe_string(ccall((:mm_last_error, libmmb200), Cstring, ())))

# batched state: N points of PlasticState, AoS = Vector{PlasticState{3,Float64,6}} uploaded as-is (14 doubles/point)
struct BatchedPlasticState{A<:AbstractMatrix{Float64}} <: AbstractMaterialState
    data::A          # AoS: 14 x N (column = one PlasticState);  SoA: N x 14
    layout::Cint
end
initial_material_state(m::Plastic, N::Integer; layout = MM_LAYOUT_AOS) =
    BatchedPlasticState(layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, 14, N) : CUDA.zeros(Float64, N, 14), layout)
get_cache(::Plastic, ::Val{:gpu}) = nothing      # the GPU path keeps the local solve in registers

"""
    material_response(m::Plastic, ε::CuMatrix{Float64}, state::BatchedPlasticState, Δt=nothing; cache, options)

N points at once; same return convention as `src/Plastic.jl:126-160`: `(σ, ∂σ∂ε, new_state)`.
Raises the reference's error if any local Newton solve did not converge.
"""
function material_response(m::Plastic, ε::CuMatrix{Float64}, state::BatchedPlasticState, Δt = nothing;
                           cache = nothing, options::Dict{Symbol,Any} = Dict{Symbol,Any}())
    aos = state.layout == MM_LAYOUT_AOS
    N = aos ? size(ε, 2) : size(ε, 1)
    σ   = similar(ε)
    tan = aos ? CUDA.zeros(Float64, 36, N) : CUDA.zeros(Float64, N, 36)
    new = similar(state.data)
    flags = CUDA.zeros(UInt8, N); iters = CUDA.zeros(UInt16, N); nfail = CUDA.zeros(Int32, 1)
    p = get(options, :nlsolve_params, Dict{Symbol,Any}())
    get(p, :method, :newton) == :newton || error("only NLsolve method = :newton exists on the GPU path (got $(p[:method]))")   # Plastic.jl:146-147
    opts = Ref(MMNewtonOpts(get(p, :ftol, 1e-8), get(p, :iterations, 1000), 0))
    rc = ccall((:mm_plastic, libmmb200), Cint,
               (Ref{MMPlastic}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble},
                CuPtr{UInt8}, CuPtr{UInt16}, CuPtr{Int32}, Ref{MMNewtonOpts}, Int64, Cint, Ptr{Cvoid}),
               Ref(MMPlastic(m)), ε, state.data, σ, tan, new, flags, iters, nfail, opts, N, state.layout,
               CUDA.stream().handle)
    check(rc, "mm_plastic")
    Array(nfail)[1] == 0 || error("Material model not converged. Could not find material state.")   # Plastic.jl:157
    return σ, tan, BatchedPlasticState(new, state.layout)
end

# host arrays (e.g. reinterpret(Float64, ::Vector{SymmetricTensor{2,3,Float64,6}}) reshaped to 6 x N): pipelined mmh_* path
function material_response(m::Plastic, ε::Matrix{Float64}, state::Matrix{Float64}, Δt = nothing; layout = MM_LAYOUT_AOS, kwargs...)
    N = layout == MM_LAYOUT_AOS ? size(ε, 2) : size(ε, 1)
    σ = similar(ε); tan = layout == MM_LAYOUT_AOS ? zeros(36, N) : zeros(N, 36); new = similar(state)
    flags = zeros(UInt8, N); iters = zeros(UInt16, N); nfail = Ref{Int32}(0)
    rc = ccall((:mmh_plastic, libmmb200), Cint,
               (Ref{MMPlastic}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{UInt16},
                Ref{Int32}, Ptr{Cvoid}, Int64, Cint),
               Ref(MMPlastic(m)), ε, state, σ, tan, new, flags, iters, nfail, C_NULL, N, layout)
    check(rc, "mmh_plastic")
    nfail[] == 0 || error("Material model not converged. Could not find material state.")
    return σ, tan, new
end

# Compact return (include/mmb200.h, mmh_plastic_compact) — the reference's own return shape, src/Plastic.jl:134-135: elastic points share
# m.Eᵉ and keep their state; each yielded point (flags & 0x01) owns one column of `rows`, in point order.  kind = :factored ->
# rows[:, k] = [gξ, u(6), n(6) | state(14)] with ∂σ∂ε = Eᵉ − gξ I_dev + u⊗n; kind = :compact -> [∂σ∂ε(36) | state(14)].
# 160 B up and 51 B + 216 B per yielded point down instead of 450 B: the host-buffer path is PCIe-bound, so this is 2.4x the throughput.
const MM_ROWS_TANGENT_STATE, MM_ROWS_FACTORED_STATE = Cint(0), Cint(1)
struct CompactResponse
    σ::Matrix{Float64}; flags::Vector{UInt8}; iters::Vector{UInt16}
    rows::Matrix{Float64}            # W x n_rows
    kind::Cint; layout::Cint
end
function material_response_compact(m::Plastic, ε::Matrix{Float64}, state::Matrix{Float64}; kind::Symbol = :factored, layout = MM_LAYOUT_AOS,
                                   capacity::Integer = (layout == MM_LAYOUT_AOS ? size(ε, 2) : size(ε, 1)))
    N = layout == MM_LAYOUT_AOS ? size(ε, 2) : size(ε, 1)
    rk = kind == :factored ? MM_ROWS_FACTORED_STATE : MM_ROWS_TANGENT_STATE
    W = Int(ccall((:mm_plastic_row_width, libmmb200), Cint, (Cint,), rk))
    σ = similar(ε); flags = zeros(UInt8, N); iters = zeros(UInt16, N); rows = zeros(W, capacity)
    nrows = Ref{Int64}(0); nfail = Ref{Int32}(0)
    rc = ccall((:mmh_plastic_compact, libmmb200), Cint,
               (Ref{MMPlastic}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}, Ptr{UInt16}, Cint, Ptr{Cdouble}, Int64, Ref{Int64}, Ptr{Cdouble},
                Ref{Int32}, Ptr{Cvoid}, Int64, Cint),
               Ref(MMPlastic(m)), ε, state, σ, flags, iters, rk, rows, capacity, nrows, C_NULL, nfail, C_NULL, N, layout)
    rc == -4 && error("material_response_compact: $(nrows[]) points yielded, capacity = $capacity")
    check(rc, "mmh_plastic_compact")
    nfail[] == 0 || error("Material model not converged. Could not find material state.")
    return CompactResponse(σ, flags, iters, rows[:, 1:nrows[]], rk, layout)
end
# dense ∂σ∂ε (36 x N) and new state (14 x N; `state` is the previous one and may be updated in place) from a CompactResponse
function expand(m::Plastic, r::CompactResponse, state::Matrix{Float64}; inplace::Bool = false)
    N = length(r.flags)
    tan = r.layout == MM_LAYOUT_AOS ? zeros(36, N) : zeros(N, 36)
    new = inplace ? state : similar(state)
    check(ccall((:mmh_plastic_expand, libmmb200), Cint,
                (Ref{MMPlastic}, Cint, Ptr{UInt8}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Cint),
                Ref(MMPlastic(m)), r.kind, r.flags, r.rows, size(r.rows, 2), state, tan, new, N, r.layout), "mmh_plastic_expand")
    return tan, new
end

# LinearElastic (src/LinearElastic.jl:57-60): the tangent is the one shared m.Eᵉ, exactly as the reference returns it
function material_response(m::LinearElastic, ε::CuMatrix{Float64}, state = initial_material_state(m), Δt = nothing;
                           layout = MM_LAYOUT_AOS, cache = nothing, options = nothing)
    N = layout == MM_LAYOUT_AOS ? size(ε, 2) : size(ε, 1)
    σ = similar(ε)
    check(ccall((:mm_linear_elastic, libmmb200), Cint,
                (Ref{MMLinearElastic}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, Int64, Cint, Ptr{Cvoid}),
                Ref(MMLinearElastic(m.E, m.ν)), ε, σ, CU_NULL, N, layout, CUDA.stream().handle), "mm_linear_elastic")
    return σ, m.Eᵉ, state
end

# NeoHook / Yeoh / StVenant through the trait route (src/traits.jl:91-107): F (9 x N) -> S (6 x N), ∂S∂C (36 x N)
_model(::NeoHook) = Cint(0); _model(::Yeoh) = Cint(1); _model(::StVenant) = Cint(2)
_hyper(m::NeoHook) = MMHyper(m.λ, m.μ, 0.0, 0.0); _hyper(m::StVenant) = MMHyper(m.λ, m.μ, 0.0, 0.0)
_hyper(m::Yeoh) = MMHyper(m.λ, m.μ, m.c₂, m.c₃)
# The reference's strain wrappers are typed on single tensors (`DeformationGradient{T<:Tensor{2}}`, src/traits.jl:2-6), so the
# batched arrays get a wrapper of their own that carries the measure as a type parameter; the outer constructors keep the
# reference's spelling: DeformationGradient(F::CuMatrix), RightCauchyGreen(C::CuMatrix), GreenLagrange(E::CuMatrix).
struct BatchedStrain{K<:MaterialModels.StrainMeasure,A<:AbstractMatrix{Float64}}
    value::A
end
MaterialModels.DeformationGradient(F::AbstractMatrix{Float64}) = BatchedStrain{DeformationGradient,typeof(F)}(F)
MaterialModels.RightCauchyGreen(C::AbstractMatrix{Float64}) = BatchedStrain{RightCauchyGreen,typeof(C)}(C)
MaterialModels.GreenLagrange(E::AbstractMatrix{Float64}) = BatchedStrain{GreenLagrange,typeof(E)}(E)
function material_response(::MaterialModels.∂S∂C, m::Union{NeoHook,Yeoh,StVenant}, F::BatchedStrain{DeformationGradient,<:CuMatrix{Float64}},
                           state, Δt::Float64 = 0.0; layout = MM_LAYOUT_AOS, cache = nothing, options = nothing)
    N = layout == MM_LAYOUT_AOS ? size(F.value, 2) : size(F.value, 1)
    S = layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, 6, N) : CUDA.zeros(Float64, N, 6)
    T = layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, 36, N) : CUDA.zeros(Float64, N, 36)
    check(ccall((:mm_hyper, libmmb200), Cint,
                (Cint, Ref{MMHyper}, CuPtr{Cdouble}, Cint, CuPtr{Cdouble}, CuPtr{Cdouble}, Int64, Cint, Ptr{Cvoid}),
                _model(m), Ref(_hyper(m)), F.value, Cint(1), S, T, N, layout, CUDA.stream().handle), "mm_hyper")
    return S, T, state
end

# XuNeedleman (XuNeedleman.jl:75-91): Δ (dim x N), cache/options positional as in the reference
function material_response(m::XuNeedleman, Δ::CuMatrix{Float64}, state = XuNeedlemanState(), Δt = nothing, cache = nothing, options = nothing;
                           layout = MM_LAYOUT_AOS)
    dim = layout == MM_LAYOUT_AOS ? size(Δ, 1) : size(Δ, 2)
    N = layout == MM_LAYOUT_AOS ? size(Δ, 2) : size(Δ, 1)
    T = similar(Δ)
    dTdΔ = layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, dim * dim, N) : CUDA.zeros(Float64, N, dim * dim)
    check(ccall((:mm_xu_needleman, libmmb200), Cint,
                (Ref{MMXuNeedleman}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, Cint, Int64, Cint, Ptr{Cvoid}),
                Ref(MMXuNeedleman(m.Φₙ, m.Φₜ, m.δₙ, m.δₜ, m.Δₙˢ)), Δ, T, dTdΔ, dim, N, layout, CUDA.stream().handle), "mm_xu_needleman")
    return T, dTdΔ, state
end

# Trait routes with another output measure (src/traits.jl:61-77): ∂S∂E (tangent_kind 1) and ∂Pᵀ∂F (tangent_kind 2,
# Pᵀ 9 x N, ∂Pᵀ∂F 81 x N) through mm_hyper_ex; strain_kind 0 = C, 1 = F, 2 = E.
_tk(::MaterialModels.∂S∂C) = Cint(0); _tk(::MaterialModels.∂S∂E) = Cint(1); _tk(::MaterialModels.∂Pᵀ∂F) = Cint(2)
_sk(::BatchedStrain{RightCauchyGreen}) = Cint(0); _sk(::BatchedStrain{DeformationGradient}) = Cint(1); _sk(::BatchedStrain{GreenLagrange}) = Cint(2)
function material_response(t::Union{MaterialModels.∂S∂E,MaterialModels.∂Pᵀ∂F}, m::Union{NeoHook,Yeoh,StVenant},
                           strain::BatchedStrain{<:Any,<:CuMatrix{Float64}}, state, Δt::Float64 = 0.0;
                           layout = MM_LAYOUT_AOS, cache = nothing, options = nothing)
    N = layout == MM_LAYOUT_AOS ? size(strain.value, 2) : size(strain.value, 1)
    ns, nt = t isa MaterialModels.∂Pᵀ∂F ? (9, 81) : (6, 36)
    S = layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, ns, N) : CUDA.zeros(Float64, N, ns)
    T = layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, nt, N) : CUDA.zeros(Float64, N, nt)
    check(ccall((:mm_hyper_ex, libmmb200), Cint,
                (Cint, Ref{MMHyper}, CuPtr{Cdouble}, Cint, Cint, CuPtr{Cdouble}, CuPtr{Cdouble}, Int64, Cint, Ptr{Cvoid}),
                _model(m), Ref(_hyper(m)), strain.value, _sk(strain), _tk(t), S, T, N, layout, CUDA.stream().handle), "mm_hyper_ex")
    return S, T, state
end

# elastic_strain_energy_density (LinearElastic.jl:63; neohook.jl:29, yeoh.jl:31, stvenant.jl:28; traits.jl:133-136): ψ[N]
function MaterialModels.elastic_strain_energy_density(m::Union{NeoHook,Yeoh,StVenant}, strain::BatchedStrain{<:Any,<:CuMatrix{Float64}};
                                                      layout = MM_LAYOUT_AOS)
    N = layout == MM_LAYOUT_AOS ? size(strain.value, 2) : size(strain.value, 1)
    ψ = CUDA.zeros(Float64, N)
    check(ccall((:mm_hyper_energy, libmmb200), Cint, (Cint, Ref{MMHyper}, CuPtr{Cdouble}, Cint, CuPtr{Cdouble}, Int64, Cint, Ptr{Cvoid}),
                _model(m), Ref(_hyper(m)), strain.value, _sk(strain), ψ, N, layout, CUDA.stream().handle), "mm_hyper_energy")
    return ψ
end
function MaterialModels.elastic_strain_energy_density(m::LinearElastic, ε::CuMatrix{Float64}; layout = MM_LAYOUT_AOS)
    N = layout == MM_LAYOUT_AOS ? size(ε, 2) : size(ε, 1)
    ψ = CUDA.zeros(Float64, N)
    check(ccall((:mm_linear_elastic_energy, libmmb200), Cint, (Ref{MMLinearElastic}, CuPtr{Cdouble}, CuPtr{Cdouble}, Int64, Cint, Ptr{Cvoid}),
                Ref(MMLinearElastic(m.E, m.ν)), ε, ψ, N, layout, CUDA.stream().handle), "mm_linear_elastic_energy")
    return ψ
end

# TransverselyIsotropic (src/transverselyisotropic.jl:92-106): the state is the unit direction a3 per point (3 x N)
struct MMTransverselyIsotropic; L_perp::Cdouble; L_par::Cdouble; M_par::Cdouble; G_perp::Cdouble; G_par::Cdouble; end
struct BatchedTransverselyIsotropicState{A<:AbstractMatrix{Float64}} <: AbstractMaterialState
    a3::A
    layout::Cint
end
function material_response(m::TransverselyIsotropic, ε::CuMatrix{Float64}, state::BatchedTransverselyIsotropicState, Δt = nothing;
                           cache = nothing, options = nothing)
    layout = state.layout
    N = layout == MM_LAYOUT_AOS ? size(ε, 2) : size(ε, 1)
    σ = similar(ε)
    E = layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, 36, N) : CUDA.zeros(Float64, N, 36)
    check(ccall((:mm_transversely_isotropic, libmmb200), Cint,
                (Ref{MMTransverselyIsotropic}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, Int64, Cint, Ptr{Cvoid}),
                Ref(MMTransverselyIsotropic(m.L⊥, m.L₌, m.M₌, m.G⊥, m.G₌)), ε, state.a3, σ, E, N, layout, CUDA.stream().handle),
          "mm_transversely_isotropic")
    return σ, E, state
end

# Dimension wrappers (src/wrappers.jl:29-79) around Plastic: dim_kind 1 UniaxialStrain, 2 PlaneStrain, 3 UniaxialStress,
# 4 PlaneStress, 5 ShellStress; reduced arrays have nc = mm_wrap_ncomp(kind) components (1, 3 or 6).
_dimkind(::UniaxialStrain) = Cint(1); _dimkind(::PlaneStrain) = Cint(2); _dimkind(::UniaxialStress) = Cint(3)
_dimkind(::PlaneStress) = Cint(4); _dimkind(::ShellStress) = Cint(5)
# Dim{d}, the generic strain-restricted state (wrappers.jl:3, dispatched with UniaxialStrain / PlaneStrain at :29-41); Dim{3} is the plain 3D call (:22)
_dimkind(::MaterialModels.Dim{1}) = Cint(1); _dimkind(::MaterialModels.Dim{2}) = Cint(2)
material_response(::MaterialModels.Dim{3}, m::MaterialModels.AbstractMaterial, Δε::CuMatrix{Float64}, args...; kwargs...) = material_response(m, Δε, args...; kwargs...)
function material_response(dim::MaterialModels.AbstractDim, m::Plastic, Δε::CuMatrix{Float64}, state::BatchedPlasticState, Δt = nothing;
                           cache = nothing, options::Dict{Symbol,Any} = Dict{Symbol,Any}())
    kind = _dimkind(dim); layout = state.layout
    nc = Int(ccall((:mm_wrap_ncomp, libmmb200), Cint, (Cint,), kind))
    N = layout == MM_LAYOUT_AOS ? size(Δε, 2) : size(Δε, 1)
    σ = similar(Δε)
    T = layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, nc * nc, N) : CUDA.zeros(Float64, N, nc * nc)
    new = similar(state.data)
    flags = CUDA.zeros(UInt8, N); iters = CUDA.zeros(UInt16, N); outer = CUDA.zeros(UInt8, N); nfail = CUDA.zeros(Int32, 1)
    inner = MMNewtonOpts(1e-8, 1000, 0)
    wrap = MMNewtonOpts(get(options, :plane_stress_tol, 1e-8), get(options, :plane_stress_max_iter, 10), 0)        # wrappers.jl:55-56
    check(ccall((:mm_wrapped_plastic, libmmb200), Cint,
                (Cint, Ref{MMPlastic}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{UInt8}, CuPtr{UInt16},
                 CuPtr{UInt8}, CuPtr{Int32}, Ref{MMNewtonOpts}, Ref{MMNewtonOpts}, Int64, Cint, Ptr{Cvoid}),
                kind, Ref(MMPlastic(m)), Δε, state.data, σ, T, new, flags, iters, outer, nfail, Ref(inner), Ref(wrap), N, layout,
                CUDA.stream().handle), "mm_wrapped_plastic")
    if Array(nfail)[1] != 0
        any(Array(flags) .& 0x40 .!= 0) && error("Not converged. Could not find plane stress state.")              # wrappers.jl:78
        error("Material model not converged. Could not find material state.")                                      # Plastic.jl:157
    end
    return σ, T, BatchedPlasticState(new, layout)
end

# The caller's quadrature loop fused around the update (include/mmb200.h mm_cells_*): what a Ferrite.jl assemble_cell! does
# around material_response, all cells at once.  dNdx: P x 3nn (shape_gradient(cv, q, a) after reinit!, P = ncell*nq, row = cell*nq + q
# -> component-major as the C side wants), dΩ: P, ue: ncell x 3nn, state: P x 14 (SoA).  Returns fe (ncell x 3nn) and the new state.
function cell_internal_forces(m::Plastic, dNdx::CuMatrix{Float64}, dΩ::CuVector{Float64}, ue::CuMatrix{Float64}, state::BatchedPlasticState;
                              tangent::Bool = false)
    state.layout == MM_LAYOUT_SOA || error("cell_internal_forces needs the SoA state layout")
    ncell, nc = size(ue); nn = nc ÷ 3; P = length(dΩ); nq = P ÷ ncell
    fe = CUDA.zeros(Float64, ncell, nc); new = similar(state.data)
    D = tangent ? CUDA.zeros(Float64, P, 36) : nothing
    flags = CUDA.zeros(UInt8, P); iters = CUDA.zeros(UInt16, P); nfail = CUDA.zeros(Int32, 1)
    check(ccall((:mm_cells_plastic, libmmb200), Cint,
                (Ref{MMPlastic}, Cint, Cint, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble},
                 CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{UInt8}, CuPtr{UInt16}, CuPtr{Int32}, Ptr{Cvoid}, Int64, Ptr{Cvoid}),
                Ref(MMPlastic(m)), nn, nq, dNdx, dΩ, ue, state.data, fe, CU_NULL #= ke: pass a (3nn)² x ncell array for BᵀDB =#, CU_NULL, tangent ? D : CU_NULL, new, flags, iters, nfail, C_NULL, ncell,
                CUDA.stream().handle), "mm_cells_plastic")
    Array(nfail)[1] == 0 || error("Material model not converged. Could not find material state.")                  # Plastic.jl:157
    return fe, BatchedPlasticState(new, MM_LAYOUT_SOA), D
end

# CrystalViscoPlastic{12} (CrystalViscoPlastic.jl:64-94): the MMCrystal struct (scalars, nslip, MS[12][9], H[12][12]) is
# filled from `m.MS` and `m.H` directly (same memory order: MS[i] column-major Tensor{2,3}; H[i][j] = m.H[i,j]).
struct MMCrystal
    E::Cdouble; nu::Cdouble; tau_y::Cdouble; H_iso::Cdouble; H_kin::Cdouble; q::Cdouble; alpha_inf::Cdouble; t_star::Cdouble; sigma_c::Cdouble; m::Cdouble
    nslip::Int32; pad::Int32
    MS::NTuple{108,Cdouble}          # MS[i][0:9], i = 0..11
    H::NTuple{144,Cdouble}           # H[i][j] = m.H[i+1, j+1]
end
MMCrystal(m::CrystalViscoPlastic{12}) = MMCrystal(m.E, m.ν, m.τ_y, m.H_iso, m.H_kin, m.q, m.α_∞, m.t_star, m.σ_c, m.m, 12, 0,
    ntuple(n -> m.MS[(n - 1) ÷ 9 + 1][(n - 1) % 9 + 1], 108), ntuple(n -> m.H[(n - 1) ÷ 12 + 1, (n - 1) % 12 + 1], 144))
struct BatchedCrystalState{A<:AbstractMatrix{Float64}} <: AbstractMaterialState
    data::A          # AoS: 42 x N (σ, κ, α, μ);  SoA: N x 42
    layout::Cint
end
function material_response(m::CrystalViscoPlastic{12}, Δε::CuMatrix{Float64}, state::BatchedCrystalState, Δt = 1.0;
                           cache = nothing, options = NamedTuple())
    layout = state.layout
    N = layout == MM_LAYOUT_AOS ? size(Δε, 2) : size(Δε, 1)
    σ = similar(Δε)
    T = layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, 36, N) : CUDA.zeros(Float64, N, 36)
    new = similar(state.data)
    active = CUDA.zeros(UInt16, N); iters = CUDA.zeros(UInt16, N); nfail = CUDA.zeros(Int32, 1)
    opts = MMNewtonOpts(get(options, :tol, 1e-8), get(options, :max_iter, 100), 0)                                  # nonlinear_solver.jl:52
    check(ccall((:mm_crystal, libmmb200), Cint,
                (Ref{MMCrystal}, CuPtr{Cdouble}, Cdouble, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{UInt16}, CuPtr{UInt16},
                 CuPtr{Int32}, Ref{MMNewtonOpts}, Int64, Cint, Ptr{Cvoid}),
                Ref(MMCrystal(m)), Δε, Float64(Δt), state.data, σ, T, new, active, iters, nfail, Ref(opts), N, layout, CUDA.stream().handle), "mm_crystal")
    Array(nfail)[1] == 0 || error("Material model not converged. Could not find material state.")                  # CrystalViscoPlastic.jl:92
    return σ, T, BatchedCrystalState(new, layout)
end

# The same wrappers around the other small-strain materials (wrappers.jl is generic in the material).
_wrap_opts(options) = MMNewtonOpts(get(options, :plane_stress_tol, 1e-8), get(options, :plane_stress_max_iter, 10), 0)      # wrappers.jl:55-56
function _reduced_outputs(kind::Cint, Δε::CuMatrix{Float64}, layout)
    nc = Int(ccall((:mm_wrap_ncomp, libmmb200), Cint, (Cint,), kind))
    N = layout == MM_LAYOUT_AOS ? size(Δε, 2) : size(Δε, 1)
    T = layout == MM_LAYOUT_AOS ? CUDA.zeros(Float64, nc * nc, N) : CUDA.zeros(Float64, N, nc * nc)
    return N, similar(Δε), T, CUDA.zeros(UInt8, N), CUDA.zeros(Int32, 1)
end
function material_response(dim::MaterialModels.AbstractDim, m::LinearElastic, Δε::CuMatrix{Float64}, state = initial_material_state(m), Δt = nothing;
                           cache = nothing, options::Dict{Symbol,Any} = Dict{Symbol,Any}(), layout = MM_LAYOUT_AOS)
    kind = _dimkind(dim)
    N, σ, T, outer, nfail = _reduced_outputs(kind, Δε, layout)
    check(ccall((:mm_wrapped_linear_elastic, libmmb200), Cint,
                (Cint, Ref{MMLinearElastic}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{UInt8}, CuPtr{Int32}, Ref{MMNewtonOpts}, Int64, Cint, Ptr{Cvoid}),
                kind, Ref(MMLinearElastic(m.E, m.ν)), Δε, σ, T, outer, nfail, Ref(_wrap_opts(options)), N, layout, CUDA.stream().handle),
          "mm_wrapped_linear_elastic")
    Array(nfail)[1] == 0 || error("Not converged. Could not find plane stress state.")                             # wrappers.jl:78
    return σ, T, state
end
function material_response(dim::MaterialModels.AbstractDim, m::TransverselyIsotropic, Δε::CuMatrix{Float64}, state::BatchedTransverselyIsotropicState,
                           Δt = nothing; cache = nothing, options::Dict{Symbol,Any} = Dict{Symbol,Any}())
    kind = _dimkind(dim); layout = state.layout
    N, σ, T, outer, nfail = _reduced_outputs(kind, Δε, layout)
    check(ccall((:mm_wrapped_transversely_isotropic, libmmb200), Cint,
                (Cint, Ref{MMTransverselyIsotropic}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{UInt8}, CuPtr{Int32},
                 Ref{MMNewtonOpts}, Int64, Cint, Ptr{Cvoid}),
                kind, Ref(MMTransverselyIsotropic(m.L⊥, m.L₌, m.M₌, m.G⊥, m.G₌)), Δε, state.a3, σ, T, outer, nfail, Ref(_wrap_opts(options)), N, layout,
                CUDA.stream().handle), "mm_wrapped_transversely_isotropic")
    Array(nfail)[1] == 0 || error("Not converged. Could not find plane stress state.")
    return σ, T, state
end
function material_response(dim::MaterialModels.AbstractDim, m::CrystalViscoPlastic{12}, Δε::CuMatrix{Float64}, state::BatchedCrystalState, Δt = 1.0;
                           cache = nothing, options::Dict{Symbol,Any} = Dict{Symbol,Any}())
    kind = _dimkind(dim); layout = state.layout
    N, σ, T, outer, nfail = _reduced_outputs(kind, Δε, layout)
    new = similar(state.data)
    active = CUDA.zeros(UInt16, N); iters = CUDA.zeros(UInt16, N)
    inner = MMNewtonOpts(get(options, :tol, 1e-8), get(options, :max_iter, 100), 0)
    check(ccall((:mm_wrapped_crystal, libmmb200), Cint,
                (Cint, Ref{MMCrystal}, CuPtr{Cdouble}, Cdouble, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{UInt16}, CuPtr{UInt16},
                 CuPtr{UInt8}, CuPtr{Int32}, Ref{MMNewtonOpts}, Ref{MMNewtonOpts}, Int64, Cint, Ptr{Cvoid}),
                kind, Ref(MMCrystal(m)), Δε, Float64(Δt), state.data, σ, T, new, active, iters, outer, nfail, Ref(inner), Ref(_wrap_opts(options)), N, layout,
                CUDA.stream().handle), "mm_wrapped_crystal")
    if Array(nfail)[1] != 0
        any(Array(active) .& 0x2000 .!= 0) && error("Not converged. Could not find plane stress state.")           # wrappers.jl:78
        error("Material model not converged. Could not find material state.")                                      # CrystalViscoPlastic.jl:92
    end
    return σ, T, BatchedCrystalState(new, layout)
end

# Small-strain trait route (src/traits.jl:91-107 with native strain SmallStrain / native tangent ∂σ∂ε, :142-143; LinearElastic.jl:25;
# test_linear_elastic.jl:32-33): both transformations are the identity (traits.jl:33, :58), so the wrapper is taken off and the native
# batched method runs.  Any other strain measure fails as the reference does (:120).
MaterialModels.SmallStrain(ε::AbstractMatrix{Float64}) = BatchedStrain{SmallStrain,typeof(ε)}(ε)
const SmallStrainMaterial = Union{LinearElastic,Plastic,TransverselyIsotropic,CrystalViscoPlastic{12}}
material_response(::MaterialModels.∂σ∂ε, m::SmallStrainMaterial, ε::BatchedStrain{SmallStrain}, args...; kwargs...) = material_response(m, ε.value, args...; kwargs...)
material_response(::MaterialModels.∂σ∂ε, m::SmallStrainMaterial, strain::BatchedStrain{K}, args...; kwargs...) where {K} =
    error("$K is not the native strain measure for $(typeof(m)):")

end # module
