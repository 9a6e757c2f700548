# bench_reference.jl — the reference's OWN CPU path on the bench workload (SURVEY.md §8(d)).
# NOT RUN in this repository: the build image has no Julia.  For anyone who has it:
#     julia -t auto --project=/path/to/MaterialModels.jl julia/bench_reference.jl [N]
# It times `material_response(::Plastic, ε, state; cache)` over N points with `Threads.@threads`, one `get_cache(m)` per thread as the
# reference requires (src/MaterialModels.jl:66-67), on the same synthetic inputs as bench.py (BASELINE.json configs[1]):
# splitmix64 uniforms, step A from the virgin state with amplitude 6.6e-4, step B (timed) with a fresh strain from A's state.
using MaterialModels, Tensors

const SEED0 = 20261019

@inline function splitmix64(x::UInt64)
    x += 0x9E3779B97F4A7C15
    x = (x ⊻ (x >> 30)) * 0xBF58476D1CE4E5B9
    x = (x ⊻ (x >> 27)) * 0x94D049BB133111EB
    return x ⊻ (x >> 31)
end
uniform(seed, i, c, nc) = Float64(splitmix64(UInt64(seed) + UInt64(i * nc + c)) >> 11) * 2.0^-53      # i, c zero-based

strain(seed, i, a) = SymmetricTensor{2,3}(ntuple(c -> a * (2uniform(seed, i, c - 1, 6) - 1), 6))      # storage order 11,21,31,22,32,33

function run(N)
    m = Plastic(E=200e3, ν=0.3, σ_y=200.0, H=50.0, r=0.5, κ_∞=13.0, α_∞=13.0)                         # test/test_plastic.jl:3
    caches = [get_cache(m) for _ in 1:Threads.nthreads()]
    states = [initial_material_state(m) for _ in 1:N]
    seedA, seedB, a = SEED0 + 2, SEED0 + 2 + 1000, 6.6e-4                                               # synth.py: c2_strain(step)
    Threads.@threads for i in 1:N                                                                       # step A (untimed)
        _, _, states[i] = material_response(m, strain(seedA, i - 1, a), states[i]; cache=caches[Threads.threadid()])
    end
    σ = Vector{SymmetricTensor{2,3,Float64,6}}(undef, N)
    D = Vector{SymmetricTensor{4,3,Float64,36}}(undef, N)
    new = similar(states)
    t = @elapsed Threads.@threads for i in 1:N                                                          # step B (timed)
        σ[i], D[i], new[i] = material_response(m, strain(seedB, i - 1, a), states[i]; cache=caches[Threads.threadid()])
    end
    println("""{"impl": "reference-julia", "metric": "material-point updates/sec, J2 plastic 3D", "value": $(N / t), "unit": "updates/s", "threads": $(Threads.nthreads()), "points": $N}""")
end

run(length(ARGS) > 0 ? parse(Int, ARGS[1]) : 1_000_000)
