# parity.jl — run wherever Julia + Microclimate.jl resolve AND a B200 is visible:
#   julia --project=/path/to/Microclimate.jl julia/parity.jl
# Runs the reference's own `solve_soil!` and the B200 engine on the reference's test problems and prints the deviation table
# (north-star tolerances: 1e-3 K soil temperature, 1e-5 m3/m3 moisture).  Until this has been run, parity vs the Julia
# reference itself is UNPINNED (DESIGN.md §5); parity vs the C++ restatement is bit-exact.
using Microclimate, Unitful
include(joinpath(@__DIR__, "MicroclimateB200.jl"))
using .MicroclimateB200

function deviation(problem)
    ref = init(problem); Microclimate.solve_solar!(ref); Microclimate.populate_weather!(ref); Microclimate.solve_soil!(ref)
    gpu = init(problem); Microclimate.solve_solar!(gpu); Microclimate.populate_weather!(gpu)
    MicroclimateB200.solve_soil_batched!([gpu])
    dT = maximum(abs.(ustrip.(u"K", ref.output.soil_temperature) .- ustrip.(u"K", gpu.output.soil_temperature)))
    dθ = maximum(abs.(ref.output.soil_moisture .- gpu.output.soil_moisture))
    return dT, dθ
end

for (name, problem) in ("example_microclimate_problem" => Microclimate.example_microclimate_problem(),)
    dT, dθ = deviation(problem)
    println(rpad(name, 36), " max|ΔT| = ", dT, " K   max|Δθ| = ", dθ, "   ", (dT <= 1e-3 && dθ <= 1e-5) ? "OK" : "ABOVE TOLERANCE")
end
