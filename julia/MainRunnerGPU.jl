# MainRunnerGPU.jl — drop-in replacement of `main_runner_tree` (MainRunner.jl:355-765) that keeps the
# signature, the keyword set, the return value and the output file, and delegates the hot path to
# libart_b200.so through `ccall` (C ABI: include/art.h).
#
# NOTE: Julia is not installed in the build image, so this file is the binding a maintainer adds; it is
# exercised line by line by its Python twin adiabatic_raytracer_b200/main_runner.py (same calls, same order).
#
#   include("MainRunnerGPU.jl")      # instead of include("MainRunner.jl") in Gen_Samples.jl:6
using NPZ

const LIBART = get(ENV, "LIBART_B200", joinpath(@__DIR__, "..", "adiabatic_raytracer_b200", "libart_b200.so"))

# mirror of `art_params` (include/art.h) — field order and types must match
mutable struct ArtParams
    Mass_a::Cdouble; Ax_g::Cdouble; theta_m::Cdouble; omega_pul::Cdouble; B0::Cdouble; rNS::Cdouble; Mass_NS::Cdouble
    bndry_lyr::Cdouble; ode_err::Cdouble; reltol::Cdouble; fix_time::Cdouble; rho_DM::Cdouble; vmean_ax::Cdouble
    prob_cutoff::Cdouble; ln_t_start::Cdouble; t0_sample::Cdouble; metric_interior_radius::Cdouble
    sampler_dt::Cdouble; dtmin::Cdouble
    flat::Cint; isotropic::Cint; melrose::Cint; thick_surface::Cint; n_maxSample::Cint; num_cutoff::Cint
    MC_nodes::Cint; max_nodes::Cint; ntimes::Cint; save_mode::Cint; interp_points_ray::Cint
    interp_points_sampler::Cint; maxiters::Cint; cross_cap::Cint; integrator::Cint; reserved_::Cint
    ArtParams() = new()
end

function art_check(rc::Cint, h::Ptr{Cvoid}=C_NULL)
    rc == 0 && return
    msg = unsafe_string(ccall((:art_last_error, LIBART), Cstring, (Ptr{Cvoid},), h))
    error("libart_b200 status $rc: $msg")
end

function main_runner_tree(Mass_a, Ax_g, θm, ωPul, B0, rNS, Mass_NS, ωProp, Ntajs, gammaF;
        flat=true, isotropic=false, melrose=false, thick_surface=true, bndry_lyr=-1,
        ode_err=1e-6, cutT=100000, fix_time=0.0, CLen_Scale=true, file_tag="",
        ntimes=1000, v_NS=[0 0 0], rho_DM=0.45, vmean_ax=220.0, saveMode=0,
        ntimes_ax=1000, dir_tag="results", n_maxSample=6, iseed=-1, num_cutoff=5,
        MC_nodes=5, max_nodes=50, prob_cutoff=1e-10, device=0)

    if iseed < 0                      # MainRunner.jl:365-372
        iseed = rand(0:100000000)
    elseif iseed == 0
        iseed = Int(rand(UInt32))
    end
    print("Using seed ", iseed, "\n")
    if saveMode < 3
        ntimes = 3                    # MainRunner.jl:379-381
    end

    p = ArtParams()
    art_check(ccall((:art_params_default, LIBART), Cint, (Ref{ArtParams},), p))
    p.Mass_a = Mass_a; p.Ax_g = Ax_g; p.theta_m = θm; p.omega_pul = ωPul; p.B0 = B0; p.rNS = rNS; p.Mass_NS = Mass_NS
    p.flat = flat; p.isotropic = isotropic; p.melrose = melrose; p.thick_surface = thick_surface
    p.bndry_lyr = bndry_lyr; p.ode_err = ode_err; p.fix_time = fix_time; p.ntimes = ntimes; p.rho_DM = rho_DM
    p.vmean_ax = vmean_ax; p.save_mode = saveMode; p.n_maxSample = n_maxSample; p.num_cutoff = num_cutoff
    p.MC_nodes = MC_nodes; p.max_nodes = max_nodes; p.prob_cutoff = prob_cutoff

    maxR = Ref{Cdouble}(0.0)
    rc = ccall((:art_max_radius, LIBART), Cint, (Ref{ArtParams}, Ref{Cdouble}), p, maxR)
    if rc == -3                       # MainRunner.jl:389-396
        print("Too small Max R.... quitting.... \n")
        print("Max radius found...\t", maxR[], "\n")
        return
    end
    art_check(rc)

    h = Ref{Ptr{Cvoid}}(C_NULL)
    art_check(ccall((:art_create, LIBART), Cint, (Ref{ArtParams}, Cint, Ref{Ptr{Cvoid}}), p, device, h))
    n_events = max(Ntajs - 1, 0)      # `while photon_trajs < desired_trajs`, photon_trajs starts at 1
    ncols = saveMode > 1 ? 31 : (saveMode > 0 ? 29 : 13)
    cap = 3 * n_events + 1024
    rows = Matrix{Float64}(undef, ncols, cap)   # column-major ncols×cap == row-major cap×ncols for C
    n_rows = Ref{Int64}(0); f_inx = Ref{Int64}(0); tot_count = Ref{Int64}(0)
    GC.@preserve rows begin
        rc = ccall((:art_run, LIBART), Cint,
                   (Ptr{Cvoid}, UInt64, Int64, Int64, Ptr{Cdouble}, Int64, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ptr{Cvoid}),
                   h[], UInt64(iseed), 1, n_events, rows, cap, n_rows, f_inx, tot_count, C_NULL)
        if rc == -4 && n_rows[] > cap   # ART_ERR_CAPACITY: the reference appends rows without bound — enlarge and fetch again
            cap = n_rows[]
            rows = Matrix{Float64}(undef, ncols, cap)
            rc = ccall((:art_fetch_rows, LIBART), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ref{Int64}), h[], rows, cap, n_rows)
        end
        art_check(rc, h[])
        art_check(ccall((:art_normalize_rows, LIBART), Cint, (Ptr{Cdouble}, Int64, Cint, Int64),
                        rows, n_rows[], ncols, f_inx[]))           # saveAll[:, 8] ./= f_inx, MainRunner.jl:749
        if saveMode > 1   # clear-text final_* / event_* files, MainRunner.jl:435-444, 593-610, 689-701
            events = Matrix{Float64}(undef, 12, max(n_events, 1)); n_ev = Ref{Int64}(0)
            art_check(ccall((:art_fetch_events, LIBART), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ref{Int64}),
                            h[], events, size(events, 2), n_ev), h[])
            # (the reference writes sln_prob before the division by f_inx; the table holds that raw value)
            art_check(ccall((:art_write_event_files, LIBART), Cint,
                            (Cstring, Cstring, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Int64),
                            dir_tag, file_tag, rows, n_rows[], events, n_ev[]))
        end
        if saveMode > 2   # one tree_<file_tag><event> file per event in saveNode's layout, MainRunner.jl:17-65, 574-578
            n_files = Ref{Int64}(0)
            art_check(ccall((:art_write_tree_files, LIBART), Cint, (Ptr{Cvoid}, Cstring, Cstring, Ref{Int64}),
                            h[], dir_tag, file_tag, n_files), h[])
        end
    end
    ccall((:art_destroy, LIBART), Cvoid, (Ptr{Cvoid},), h[])

    saveAll = permutedims(rows[1:min(ncols, 29), 1:n_rows[]])                     # n_rows × ncols, as the reference builds it
    fileN = dir_tag*"/npy/tree_"                                    # MainRunner.jl:750-761 verbatim
    fileN *= "MassAx_"*string(Mass_a)*"_AxionG_"*string(Ax_g)
    fileN *="_ThetaM_"*string(θm)*"_rotPulsar_"*string(ωPul)*"_B0_"*string(B0)
    fileN *= "_Ax_trajs_"*string(Ntajs)
    fileN *= "_N_Times_"*string(ntimes);
    fileN *= "_num_cutoff_"*string(num_cutoff)
    fileN *= "_MC_nodes_"*string(MC_nodes)
    fileN *= "_max_nodes_"*string(max_nodes)
    fileN *= "_"*file_tag*".npy"
    npzwrite(fileN, saveAll)
    return tot_count[]
end


# One job over several GPUs, one Julia process per GPU (replaces runner_GR_tasks.sh:22-28 + Combine_Files.py): the events
# 1 … Ntajs−1 of ONE seeded job are split in contiguous ranges over the ranks, every rank writes its own
# tree_…_<file_tag><rank>.npy, and column 8 is normalised by the ALL-REDUCED f_inx (MainRunner.jl:749).  comm_id: the 128
# bytes of art_comm_unique_id from rank 0, passed on by any means (a file, MPI, sockets).  Python mirror with the same
# semantics: adiabatic_raytracer_b200/main_runner.py, main_runner_tree_multi (verified on 2 and 8 GPUs).
function main_runner_tree_multi(p::ArtParams, Ntajs, rank, world, comm_id::Vector{UInt8}; iseed=1769, device=rank, n_costheta=16, n_phi=32)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    art_check(ccall((:art_create, LIBART), Cint, (Ref{ArtParams}, Cint, Ref{Ptr{Cvoid}}), p, device, h))
    art_check(ccall((:art_hist_config, LIBART), Cint, (Ptr{Cvoid}, Int32, Int32), h[], n_costheta, n_phi), h[])
    world > 1 && art_check(ccall((:art_comm_init, LIBART), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), h[], world, rank, comm_id), h[])
    n_total = max(Ntajs - 1, 0)
    base, rem = divrem(n_total, world)
    first = 1 + rank * base + min(rank, rem); count = base + (rank < rem ? 1 : 0)
    n_rows = Ref{Int64}(0); f_inx = Ref{Int64}(0); tot_count = Ref{Int64}(0)
    art_check(ccall((:art_run_device, LIBART), Cint, (Ptr{Cvoid}, UInt64, Int64, Int64, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ptr{Cvoid}),
                    h[], UInt64(iseed), first, count, n_rows, f_inx, tot_count, C_NULL), h[])
    art_check(ccall((:art_comm_allreduce_hist, LIBART), Cint, (Ptr{Cvoid},), h[]), h[])       # the only collective
    len = ccall((:art_hist_len, LIBART), Int64, (Ptr{Cvoid},), h[])
    hist = zeros(len)
    art_check(ccall((:art_hist_get_global, LIBART), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64), h[], hist, len), h[])
    f_inx_total = Int64(hist[end-7])                                                          # trailing counters: f_inx, n_rows, n_events, …
    ncols = p.save_mode > 1 ? 31 : (p.save_mode > 0 ? 29 : 13)
    rows = Matrix{Float64}(undef, ncols, max(n_rows[], 1))
    art_check(ccall((:art_fetch_rows, LIBART), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Ref{Int64}), h[], rows, size(rows, 2), n_rows), h[])
    art_check(ccall((:art_normalize_rows, LIBART), Cint, (Ptr{Cdouble}, Int64, Cint, Int64), rows, n_rows[], ncols, f_inx_total))
    ccall((:art_destroy, LIBART), Cvoid, (Ptr{Cvoid},), h[])
    return permutedims(rows[1:min(ncols, 29), 1:n_rows[]]), hist, tot_count[]                 # caller: npzwrite(<name>_<tag><rank>.npy, …)
end
