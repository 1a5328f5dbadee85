"""ctypes mirror of ``art_params`` / ``art_stats`` (include/art.h).

The fields are the arguments of the reference's ``main_runner_tree`` (MainRunner.jl:355-363)
plus the fixed block of Gen_Samples.jl:158-174; defaults are the reference's.
"""
import ctypes as C
import math


class ArtParams(C.Structure):
    _fields_ = [
        ("Mass_a", C.c_double),
        ("Ax_g", C.c_double),
        ("theta_m", C.c_double),
        ("omega_pul", C.c_double),
        ("B0", C.c_double),
        ("rNS", C.c_double),
        ("Mass_NS", C.c_double),
        ("bndry_lyr", C.c_double),
        ("ode_err", C.c_double),
        ("reltol", C.c_double),
        ("fix_time", C.c_double),
        ("rho_DM", C.c_double),
        ("vmean_ax", C.c_double),
        ("prob_cutoff", C.c_double),
        ("ln_t_start", C.c_double),
        ("t0_sample", C.c_double),
        ("metric_interior_radius", C.c_double),
        ("sampler_dt", C.c_double),
        ("dtmin", C.c_double),
        ("flat", C.c_int32),
        ("isotropic", C.c_int32),
        ("melrose", C.c_int32),
        ("thick_surface", C.c_int32),
        ("n_maxSample", C.c_int32),
        ("num_cutoff", C.c_int32),
        ("MC_nodes", C.c_int32),
        ("max_nodes", C.c_int32),
        ("ntimes", C.c_int32),
        ("save_mode", C.c_int32),
        ("interp_points_ray", C.c_int32),
        ("interp_points_sampler", C.c_int32),
        ("maxiters", C.c_int32),
        ("cross_cap", C.c_int32),
        ("integrator", C.c_int32),
        ("reserved_", C.c_int32),
    ]

    @classmethod
    def default(cls, **overrides):
        """Reference defaults: Gen_Samples.jl:18-131 (flags), :158-174 (fixed block)."""
        p = cls(
            Mass_a=1e-5, Ax_g=1e-12, theta_m=0.0, omega_pul=1.0, B0=1e14, rNS=10.0, Mass_NS=1.0,
            bndry_lyr=-1.0, ode_err=1e-6, reltol=1e-7, fix_time=0.0, rho_DM=0.45, vmean_ax=220.0,
            prob_cutoff=1e-10, ln_t_start=-30.0, t0_sample=0.0, metric_interior_radius=10.0,
            sampler_dt=0.5, dtmin=1e-13,
            flat=0, isotropic=0, melrose=1, thick_surface=1, n_maxSample=6, num_cutoff=5,
            MC_nodes=5, max_nodes=50, ntimes=3, save_mode=0, interp_points_ray=50,
            interp_points_sampler=20, maxiters=100000, cross_cap=16, integrator=0, reserved_=0,
        )
        for k, v in overrides.items():
            if not hasattr(p, k):
                raise AttributeError(f"art_params has no field {k!r}")
            setattr(p, k, v)
        return p

    def copy(self, **overrides):
        q = type(self).from_buffer_copy(self)
        for k, v in overrides.items():
            if not hasattr(q, k):
                raise AttributeError(f"art_params has no field {k!r}")
            setattr(q, k, v)
        return q

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class ArtStats(C.Structure):
    _fields_ = [
        ("n_events", C.c_int64),
        ("n_attempts", C.c_int64),
        ("n_sampler_cond", C.c_int64),
        ("n_rays", C.c_int64),
        ("n_rays_photon", C.c_int64),
        ("n_acc", C.c_int64),
        ("n_rej", C.c_int64),
        ("n_rhs", C.c_int64),
        ("n_cond", C.c_int64),
        ("n_final_photon", C.c_int64),
        ("n_final_axion", C.c_int64),
        ("n_status", C.c_int64 * 5),
        ("ms_sampler", C.c_double),
        ("ms_weights", C.c_double),
        ("ms_tracer", C.c_double),
        ("ms_tree", C.c_double),
        ("ms_total", C.c_double),
        ("n_launches", C.c_int64),
        ("n_cond_eval", C.c_int64),
        ("n_scanned_steps", C.c_int64),
        ("n_pool_blocks", C.c_int64),
        ("ms_k_stages", C.c_double),
        ("ms_k_scan", C.c_double),
        ("ms_k_detect", C.c_double),
    ]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if name == "n_status" else v
        return d


# status / species codes (include/art.h)
ART_OK = 0
ART_ERR_ARG = -1
ART_ERR_CUDA = -2
ART_ERR_NO_SURFACE = -3
ART_ERR_CAPACITY = -4
ART_ERR_NCCL = -5
ART_ERR_IO = -6
RAY_END, RAY_HIT_NS, RAY_CROSSING, RAY_MAXITERS, RAY_CROSS_OVERFLOW = range(5)
SPECIES_AXION, SPECIES_PHOTON = 0, 1
CROSS_REC = 8
ROW_COLS_SHORT, ROW_COLS_LONG = 13, 29

# Constants.jl:3-5
c_km = 2.99792e5
hbar = 6.582119e-16
GNew = 1.32712e11


def ln_t_end(p: ArtParams) -> float:
    """MainRunner.jl:412."""
    return math.log(1.0 / p.omega_pul)
