# make_julia_fixtures.jl -- pins the CPU oracle (oracle/cpm_oracle.c) and the CUDA path to REAL Julia output.
#
#     julia tests/golden/make_julia_fixtures.jl /path/to/CarParkingMaps [out_dir = tests/golden/julia]
#
# Julia is not installed in the build image, so this script cannot be run there; it is the operators' one command.
# Once `tests/golden/julia/manifest.json` exists, tests/test_julia_fixtures.py compares the oracle (CPU) and libcpm.so
# (GPU, through the C ABI) against these files and the skip reason "parity unpinned" disappears.
#
# What it does
#   * `include`s the reference's seven hot-path files UNMODIFIED (src/createpdrive.jl, createpdestin.jl,
#     initializestates.jl, solveinitialvalueproblem.jl, resampling.jl, averagedrivingtime.jl, correctparameters.jl)
#     into a sandbox module that defines the globals they read (T, p_min, p_max, e_drive, e_dest, cars_per_zone;
#     main.jl:37-42) and its own `rand`:
#       - `rand()` pops the next 52-bit uniform of a splitmix64 stream -- the very generator of the oracle's "stream"
#         mode (oracle/cpm_oracle.c stream_next), so the oracle run with the same seed consumes the same variates in
#         the same order (src/resampling.jl:13,29,68,74) and must reproduce every index bit for bit;
#       - `rand(Truncated(Normal(mu, sigma), lo, hi))` takes ONE uniform from the same stream and returns the
#         truncated normal's quantile at it, through Distributions.jl when that package is installed (pinning its
#         arithmetic), else through SpecialFunctions' erf / erfinv.
#   * builds the synthetic cities from the same counter hash as carparkingmaps_b200/synth.py (ported below; the inputs
#     are written out too, so the test also pins the port), runs main.jl:82-98 on them and writes every array as raw
#     little-endian column-major binary plus a JSON manifest.
#
# Nothing here is part of the product; it is test infrastructure like oracle/.

const REF = length(ARGS) >= 1 ? ARGS[1] : error("usage: julia make_julia_fixtures.jl /path/to/CarParkingMaps [out_dir]")
const OUT = length(ARGS) >= 2 ? ARGS[2] : joinpath(@__DIR__, "julia")
mkpath(OUT)

# ---------------------------------------------------------------------------------------------------------------
# synthetic city: port of carparkingmaps_b200/synth.py (same operation order => same bits)
# ---------------------------------------------------------------------------------------------------------------
function mix(z::UInt64)
    z += 0x9E3779B97F4A7C15
    z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
    z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
    return z ⊻ (z >> 31)
end
function hashu(seed, tag, a, b, c)
    h = mix(UInt64(seed) * 0x00000100000001B3 + UInt64(tag))
    h = mix(h ⊻ UInt64(a)); h = mix(h ⊻ UInt64(b)); h = mix(h ⊻ UInt64(c))
    return Float64(h >> 11) * 2.0^-53
end
function bump(x)
    q = x / 2.5
    r = 1.0 - q * q
    return r > 0.0 ? r * r : 0.0
end
function synth_city(N, T, seed, density)
    x = [40.0 * hashu(seed, 1, i - 1, 0, 0) for i in 1:N]
    y = [40.0 * hashu(seed, 2, i - 1, 0, 0) for i in 1:N]
    d = zeros(N, N)
    dm = zeros(N, N, T, 2)
    for j in 1:N, i in 1:N
        dx = x[i] - x[j]; dy = y[i] - y[j]
        v = sqrt(dx * dx + dy * dy)
        d[i, j] = i == j ? 1.0 : (v < 0.05 ? 0.05 : v)
    end
    for j in 1:N, i in 1:N
        i == j && continue
        v0 = 30.0 + 30.0 * hashu(seed, 3, i - 1, j - 1, 0)
        amp = 0.3 + 0.9 * hashu(seed, 4, i - 1, j - 1, 0)
        phase = -1.0 + 2.0 * hashu(seed, 5, i - 1, j - 1, 0)
        base = (d[i, j] * 3600.0) / v0
        base = base < 60.0 ? 60.0 : base
        base = base > 1500.0 ? 1500.0 : base
        for t in 1:T
            present = density < 1.0 ? hashu(seed, 8, i - 1, j - 1, t - 1) < density : true
            h = Float64(((t - 1) % 24) + 1)
            c = 1.0 + amp * (bump(h - 8.0 - phase) + bump(h - 17.5 - phase)) + 0.15 * hashu(seed, 9, i - 1, j - 1, t - 1)
            m = base * c
            sd = m * (0.05 + 0.55 * hashu(seed, 6, i - 1, j - 1, t - 1))
            sd = hashu(seed, 7, i - 1, j - 1, t - 1) < 0.02 ? 0.0 : sd
            if present
                dm[i, j, t, 1] = m; dm[i, j, t, 2] = sd
            end
        end
    end
    return dm, d
end
function edge_case_city(T)   # synth.py edge_case_city: the reference's corner cases (SURVEY 8a)
    dm, d = synth_city(6, T, 1234, 1.0)
    dm[2, :, 3, :] .= 0.0
    dm[3, 5, :, 1] .= 777.0
    dm[4, :, :, :] .= 0.0
    dm[5, 6, :, 2] .= 0.0
    return dm, d
end

# ---------------------------------------------------------------------------------------------------------------
# sandbox: the reference's files, unmodified, with an injected variate stream
# ---------------------------------------------------------------------------------------------------------------
module RefSandbox
    const HAVE_DISTRIBUTIONS = try
        @eval import Distributions
        true
    catch
        false
    end
    if !HAVE_DISTRIBUTIONS
        @eval import SpecialFunctions
    end
    const STATE = Ref{UInt64}(0)
    const DRAWS = Ref{Int}(0)
    seed!(s) = (STATE[] = UInt64(s); DRAWS[] = 0; nothing)
    # oracle/cpm_oracle.c stream_next: splitmix64, top 52 bits
    function rand()
        STATE[] += 0x9E3779B97F4A7C15
        z = STATE[]
        z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
        z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
        z = z ⊻ (z >> 31)
        DRAWS[] += 1
        return Float64(z >> 12) * 2.0^-52
    end
    struct Normal; mu::Float64; sigma::Float64; end
    struct Truncated; d::Normal; lo::Float64; hi::Float64; end
    function rand(t::Truncated)
        u = rand()
        if HAVE_DISTRIBUTIONS
            return Distributions.quantile(Distributions.truncated(Distributions.Normal(t.d.mu, t.d.sigma), t.lo, t.hi), u)
        else
            cdf(x) = 0.5 * (1.0 + SpecialFunctions.erf((x - t.d.mu) / (t.d.sigma * sqrt(2.0))))
            a = cdf(t.lo); b = cdf(t.hi)
            x = t.d.mu + t.d.sigma * sqrt(2.0) * SpecialFunctions.erfinv(2.0 * (a + u * (b - a)) - 1.0)
            return min(max(x, t.lo), t.hi)
        end
    end
    # globals the reference functions read at call time (main.jl:37-42)
    T = 24; p_min = 0.1; p_max = 0.9; e_drive = 0.5; e_dest = 2; cars_per_zone = 1000
end

for f in ("createpdrive", "createpdestin", "initializestates", "solveinitialvalueproblem", "resampling",
          "averagedrivingtime", "correctparameters")
    Base.include(RefSandbox, joinpath(REF, "src", f * ".jl"))
end

setglobal(name, v) = Core.eval(RefSandbox, :($name = $v))

# ---------------------------------------------------------------------------------------------------------------
# cases
# ---------------------------------------------------------------------------------------------------------------
function dumparr(name, key, a)
    path = joinpath(OUT, "$(name)_$(key).bin")
    open(path, "w") do io; write(io, a); end
    return "\"$key\": {\"file\": \"$(name)_$(key).bin\", \"dtype\": \"$(eltype(a) == Int64 ? "int64" : "float64")\", \"shape\": [$(join(size(a), ", "))]}"
end

cases = [  # name, N, cars_per_zone, city seed, density, e_drive, e_dest, p_min, p_max
    ("edge6", 6, 50, -1, 1.0, 0.5, 2, 0.1, 0.9),
    ("n33", 33, 40, 1, 0.5, 0.5, 2, 0.1, 0.9),
    ("n100", 100, 100, 2, 1.0, 0.5, 2, 0.1, 0.9),
    ("n40pow", 40, 25, 3, 0.09, 0.6568789120171143, 0.7148157904492584, 0.0, 0.475918),   # non-integral exponents (README.md:1415, :2408)
]
entries = String[]
for (k, (name, N, cpz, cseed, density, e_drive, e_dest, p_min, p_max)) in enumerate(cases)
    T = 24
    dm, dist = cseed < 0 ? edge_case_city(T) : synth_city(N, T, cseed, density)
    setglobal(:T, T); setglobal(:p_min, p_min); setglobal(:p_max, p_max); setglobal(:e_drive, e_drive)
    setglobal(:e_dest, e_dest); setglobal(:cars_per_zone, cpz)
    C = N * cpz
    p_drive = Base.invokelatest(RefSandbox.createpdrive, dm, dist, N)                                   # main.jl:82
    p_dest = Base.invokelatest(RefSandbox.createpdestin, dm, N)                                         # main.jl:85
    state, trans = Base.invokelatest(RefSandbox.initializestates, C)                                    # main.jl:88
    ok = true
    parts = String[dumparr(name, "datamatrix", dm), dumparr(name, "distance_matrix_km", dist), dumparr(name, "p_drive", p_drive),
                   dumparr(name, "p_dest", p_dest)]
    ivp_seed = 1000 + k; day_seed = 2000 + k
    draws_ivp = 0; draws_day = 0; A = 0.0
    try
        RefSandbox.seed!(ivp_seed)
        initial_state = Base.invokelatest(RefSandbox.solveinitialvalueproblem, state, trans, p_drive, p_dest, C, N)   # main.jl:91
        draws_ivp = RefSandbox.DRAWS[]
        push!(parts, dumparr(name, "initial_state", initial_state), dumparr(name, "state_ivp", copy(state)), dumparr(name, "trans_ivp", copy(trans)))
        state[:, 1] = initial_state                                                                    # main.jl:92
        RefSandbox.seed!(day_seed)
        state, trans = Base.invokelatest(RefSandbox.resampling, state, trans, C, N, p_drive, p_dest, dm, dist)       # main.jl:95
        draws_day = RefSandbox.DRAWS[]
        A = Base.invokelatest(RefSandbox.averagedrivingtime, C, 0.0, trans)                            # main.jl:98
        push!(parts, dumparr(name, "state", state), dumparr(name, "trans", trans))
    catch err
        # the reference raises BoundsError after a draw that finds no bin (src/resampling.jl:32); the edge-case city
        # provokes that on purpose -- its tables are still pinned, its trajectories are not written
        ok = false
        println("case $name: sampling stopped with $(typeof(err)) (expected for NaN rows); tables only")
    end
    push!(entries, "\"$name\": {\"N\": $N, \"cars_per_zone\": $cpz, \"T\": $T, \"city_seed\": $cseed, \"density\": $density, " *
                   "\"e_drive\": $e_drive, \"e_dest\": $e_dest, \"p_min\": $p_min, \"p_max\": $p_max, \"sampled\": $ok, " *
                   "\"ivp_seed\": $ivp_seed, \"day_seed\": $day_seed, \"draws_ivp\": $draws_ivp, \"draws_day\": $draws_day, " *
                   "\"A_drive\": \"$(repr(A))\", \"arrays\": {" * join(parts, ", ") * "}}")
end
cp = [Base.invokelatest(RefSandbox.correctparameters, a, b, 0.1, 0.9) for (a, b) in ((-0.2, 1.3), (0.95, 0.05), (0.3, 0.7))]
open(joinpath(OUT, "manifest.json"), "w") do io
    println(io, "{\"julia_version\": \"$(VERSION)\", \"distributions_jl\": $(RefSandbox.HAVE_DISTRIBUTIONS), ")
    println(io, " \"correctparameters\": [" * join(["[$(c[1]), $(c[2])]" for c in cp], ", ") * "],")
    println(io, " \"cases\": {" * join(entries, ",\n  ") * "}}")
end
println("wrote $(length(entries)) cases to $OUT")
