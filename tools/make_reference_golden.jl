# tools/make_reference_golden.jl — golden vectors from the REAL reference, for a machine that has Julia.
#
# The build container has no Julia and no network, so the oracle is pinned by sympy/mpmath/scipy fixtures only
# ("parity unpinned", DESIGN.md §2).  This script closes that gap wherever Julia (>= 1.6) and the reference's
# dependencies (OrdinaryDiffEq, DiffEqBase, ForwardDiff, NLsolve, NPZ, SpecialFunctions, JSON) are available:
#
#     julia tools/make_reference_golden.jl /path/to/Adiabatic_RayTracer/src tests/golden
#
# It reads the inputs of the committed fixtures (tests/golden/rays.json, conv_prob.json, sampler_thick.json), calls the
# reference's own functions on them and writes tests/golden/julia_rays.json, julia_conv_prob.json, julia_sampler.json in
# the same layout.  tests/test_julia_golden.py compares the oracle (CPU) and the CUDA library (GPU) with those files
# when they exist and is skipped otherwise.  NOT RUN in the build container (no Julia): treat as untested glue.
#
# Reference entry points used (file:line of /root/reference/src):
#   RT.propagate            RayTracer.jl:171   (Vern6, reltol 1e-7, abstol ode_err, ContinuousCallback interp_points=50)
#   RT.find_samples_new     RayTracer.jl:1480  (pre_randomized hook, :1482-1518 — NB: that branch uses the 1/r radius law
#                                               rRND = u6·maxR, so sqrt(u6) is passed to reproduce the default law)
#   RT.k_norm_Cart :643, RT.k_sphere :983, RT.dwp_ds :1327 (cos_w), RT.g_det :734, get_Prob_nonAD MainRunner.jl:67
using JSON
using Random

src = ARGS[1]
out = length(ARGS) > 1 ? ARGS[2] : "tests/golden"
include(joinpath(src, "RayTracer.jl"))
include(joinpath(src, "MainRunner.jl"))   # defines RT = RayTracerGR, get_Prob_nonAD, get_tree, main_runner_tree
import .Constants: c_km, hbar, GNew

# ------------------------------------------------------------------------------------------------ rays (propagate)
rays = JSON.parsefile(joinpath(out, "rays.json"))
res = []
for ray in rays
    par = ray["par"]
    θm, ωPul, B0, rNS, MNS, ma = par["thm"], par["w"], par["B0"], par["rNS"], par["MNS"], par["ma"]
    global Mass_NS = MNS                                  # get_Prob_nonAD reads this global (MainRunner.jl:75)
    x0 = reshape(Float64.(ray["x0"]), 1, 3)
    k0 = reshape(Float64.(ray["khat"]), 1, 3)
    erg = Float64(ray["erg"])
    is_axion = ray["species"] == 0
    NumerPass = [-30.0, log(1.0 / ωPul), 1e-6]            # MainRunner.jl:411-413
    if is_axion
        Mvars = [θm, ωPul, B0, rNS, [1.0, 1.0], zeros(1), MNS, [erg], false, false, true, ma, -1]
        r = RT.propagate(x0, k0, 3, Mvars, NumerPass, RT.func_axion!, true, true, ma, 100000, Float64(ray["dOmega"]))
    else
        Mvars = [θm, ωPul, B0, rNS, [1.0, 1.0], zeros(1), MNS, ma, [erg], false, false, true, -1]
        r = RT.propagate(x0, k0, 3, Mvars, NumerPass, RT.func!, true, false, ma, 100000, Float64(ray["dOmega"]))
    end
    x_e, k_e, t_e, err_e, cut_short, xc, yc, zc, kxc, kyc, kzc, tc, Δωc, times = r
    push!(res, Dict("species" => ray["species"], "par" => par, "erg" => erg, "x0" => ray["x0"], "khat" => ray["khat"],
                    "dOmega" => ray["dOmega"], "x_f" => x_e[1, :, end], "k_f" => k_e[1, :, end], "u7_f" => t_e[1, end],
                    "crossings" => [Dict("x" => [xc[j], yc[j], zc[j]], "k" => [kxc[j], kyc[j], kzc[j]], "t" => tc[j],
                                         "dw" => Δωc[j]) for j in 1:length(xc)],
                    "lnt_saved" => times, "tolerances" => Dict("reltol" => 1e-7, "abstol" => 1e-6)))
end
open(joinpath(out, "julia_rays.json"), "w") do f
    JSON.print(f, res)
end

# ------------------------------------------------------------------------------ weights and conversion probability
cps = JSON.parsefile(joinpath(out, "conv_prob.json"))
res = []
for c in cps
    par = c["par"]
    θm, ωPul, B0, rNS, MNS, ma, Axg = par["thm"], par["w"], par["B0"], par["rNS"], par["MNS"], par["ma"], par["Ax_g"]
    global Mass_NS = MNS
    x = reshape(Float64.(c["x"]), 1, 3)
    vloc = reshape(Float64.(c["v_loc"]), 1, 3)
    vinf = reshape(Float64.(c["v_inf"]), 1, 3)
    vIfty_mag = sqrt.(sum(vinf .^ 2, dims=2))
    gammaA = 1 ./ sqrt.(1.0 .- vIfty_mag .^ 2)
    erg_inf = ma .* sqrt.(1 .+ (vIfty_mag .* gammaA) .^ 2)                                    # MainRunner.jl:523-524
    rmag = sqrt.(sum(x .^ 2, dims=2))
    k_init = RT.k_norm_Cart(x, vloc, 0.0, erg_inf, θm, ωPul, abs(B0), rNS, MNS, ma, melrose=true, isotropic=false,
                            flat=false, ax_fix=true)                                          # MainRunner.jl:529
    ksphere = RT.k_sphere(x, k_init, θm, ωPul, abs(B0), rNS, zeros(1), MNS, ma, false, bndry_lyr=-1)
    erg_ax = erg_inf ./ sqrt.(1.0 .- 2 * GNew .* MNS ./ rmag ./ c_km .^ 2)
    Mvars = [θm, ωPul, abs(B0), rNS, [1.0, 1.0], zeros(1), MNS, ma, false, false, erg_ax, -1]
    dw = RT.dwp_ds(x, ksphere, Mvars)
    cos_w = dw[4]
    x0_pl = [rmag acos.(x[:, 3] ./ rmag) atan.(x[:, 2], x[:, 1])]
    jac = RT.g_det(x0_pl, zeros(1), θm, ωPul, abs(B0), rNS, MNS, ma; flat=false, bndry_lyr=-1)
    P = get_Prob_nonAD(x, k_init, ma, Axg, θm, ωPul, B0, rNS, erg_inf, vIfty_mag, false, false, -1)
    push!(res, Dict("par" => par, "x" => c["x"], "v_loc" => c["v_loc"], "v_inf" => c["v_inf"], "k_init" => k_init[1, :],
                    "erg_inf" => erg_inf[1], "cos_w" => cos_w[1], "jac_gr" => jac[1], "P_nonAD" => P[1]))
end
open(joinpath(out, "julia_conv_prob.json"), "w") do f
    JSON.print(f, res)
end

# -------------------------------------------------------------------------------------------------------- sampler
smp = JSON.parsefile(joinpath(out, "sampler_thick.json"))
res = []
Random.seed!(1769)
for s in smp
    par = s["par"]
    U = Float64.(s["U"])
    pre = [U[1], U[2], U[3], U[4], U[5], sqrt(U[6])]       # (the pre_randomized branch multiplies maxR by u6 itself)
    xs = []
    for rep in 1:12                                        # the crossing index and v_infinity are drawn by rand(): repeat
        xpos, rRND, numV, weights, vloc, vinf = RT.find_samples_new(s["maxR"], par["thm"], par["w"], par["B0"], par["rNS"], par["ma"],
                                                                     par["MNS"]; n_max=par["n_max"], batchsize=1, thick_surface=true,
                                                                     iso=false, melrose=true, pre_randomized=pre, flat=false)
        if numV > 0
            push!(xs, Dict("x" => xpos[1, :], "v_loc" => vloc[1, :], "v_inf" => vinf[1, :]))
        end
    end
    push!(res, Dict("par" => par, "maxR" => s["maxR"], "U" => s["U"], "accepted_draws" => xs))
end
open(joinpath(out, "julia_sampler.json"), "w") do f
    JSON.print(f, res)
end
println("wrote julia_rays.json, julia_conv_prob.json, julia_sampler.json to ", out)
