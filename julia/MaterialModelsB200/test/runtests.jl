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
