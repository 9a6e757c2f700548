# NOT EXECUTED in this repository (no Julia in the build image).  What a maintainer would run on a B200 box with
# LIBMMB200=/path/to/libmmb200.so: the batched GPU methods against the reference's own per-point methods on the same inputs
# — the same comparison tests/test_gpu_parity.py makes against the oracle.
using Test, Tensors, CUDA, MaterialModels, MaterialModelsB200
const B = MaterialModelsB200

@testset "J2 Plastic: batched GPU path == per-point reference" begin
    m = Plastic(E=200e3, ν=0.3, σ_y=200.0, H=50.0, r=0.5, κ_∞=13.0, α_∞=13.0)          # test/test_plastic.jl:3
    N = 10_000
    strains = [SymmetricTensor{2,3}((i, j) -> 6.6e-4 * (2rand() - 1)) for _ in 1:N]
    states = [initial_material_state(m) for _ in 1:N]
    cache = get_cache(m)
    ref = [material_response(m, strains[i], states[i]; cache=cache) for i in 1:N]     # the caller-owned loop of the reference
    ε = CuArray(reshape(reinterpret(Float64, strains), 6, N))                          # AoS: zero-copy view of the assembler's array
    st = initial_material_state(m, N)                                                  # BatchedPlasticState, 14 x N zeros
    σ, D, st2 = material_response(m, ε, st)
    σh, Dh = Array(σ), Array(D)
    for i in 1:N
        @test σh[:, i] ≈ collect(reinterpret(Float64, [ref[i][1]])) rtol = 1e-12
        @test Dh[:, i] ≈ collect(reinterpret(Float64, [ref[i][2]])) rtol = 1e-12
    end
end

@testset "NeoHook from F through the trait route" begin
    m = NeoHook(λ=1.0, μ=1.0)
    N = 1000
    Fs = [one(Tensor{2,3}) + 0.2 * Tensor{2,3}((i, j) -> 2rand() - 1) for _ in 1:N]
    F = CuArray(reshape(reinterpret(Float64, Fs), 9, N))
    S, dSdC, _ = material_response(∂S∂C(), m, DeformationGradient(F), initial_material_state(m))
    for i in 1:10:N
        Sr, Tr, _ = material_response(∂S∂C(), m, DeformationGradient(Fs[i]), initial_material_state(m))
        @test Array(S)[:, i] ≈ collect(reinterpret(Float64, [Sr.value])) rtol = 1e-12
    end
end

@testset "J2 Plastic, host arrays: compact return == dense return (src/Plastic.jl:134-135 shape)" begin
    m = Plastic(E=200e3, ν=0.3, σ_y=200.0, H=50.0, r=0.5, κ_∞=13.0, α_∞=13.0)
    N = 50_000
    ε = 6.6e-4 .* (2 .* rand(6, N) .- 1)
    st = zeros(14, N)
    σ, D, st2 = material_response(m, ε, st)                                  # mmh_plastic (dense)
    r = B.material_response_compact(m, ε, st; kind = :factored)              # mmh_plastic_compact
    @test r.σ == σ
    @test count(!iszero, r.flags .& 0x01) == size(r.rows, 2)
    Dx, stx = B.expand(m, r, st)                                             # mmh_plastic_expand
    @test Dx == D && stx == st2                                              # bit-identical
    Ee = reshape(collect(reinterpret(Float64, [m.Eᵉ])), 36)
    @test all(Dx[:, i] == Ee for i in findall(iszero, r.flags .& 0x01))      # elastic points: the one stored Eᵉ
end

@testset "small-strain trait route and Dim{d} (src/traits.jl:91-148, src/wrappers.jl:3)" begin
    m = LinearElastic(E=200e3, ν=0.3)
    N = 1000
    ε = CuArray(2e-3 .* (2 .* rand(6, N) .- 1))
    σ1, _, _ = material_response(m, ε)
    σ2, _, _ = material_response(∂σ∂ε(), m, SmallStrain(ε))
    @test Array(σ1) == Array(σ2)
    @test_throws ErrorException material_response(∂σ∂ε(), m, DeformationGradient(CuArray(rand(9, N))))     # traits.jl:120
    p = Plastic(E=200e3, ν=0.3, σ_y=200.0, H=50.0, r=0.5, κ_∞=13.0, α_∞=13.0)
    st = initial_material_state(p, N)
    e2 = CuArray(1.5e-3 .* (2 .* rand(3, N) .- 1))
    a = material_response(PlaneStrain(), p, e2, st)
    b = material_response(MaterialModels.Dim{2}(), p, e2, st)
    @test Array(a[1]) == Array(b[1]) && Array(a[2]) == Array(b[2])
end
