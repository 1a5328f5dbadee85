"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/art.h declares,
and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from adiabatic_raytracer_b200.params import ArtParams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from adiabatic_raytracer_b200 import build, lib
    build.build()
    return lib.load()


def test_every_declared_symbol_is_exported(L):
    from adiabatic_raytracer_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "art.h")).read()
    declared = set(re.findall(r"\b(art_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"art_handle", "art_params", "art_stats"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/art.h but not exported"
        assert name in lib.SIGNATURES, f"{name} has no ctypes signature"
    assert set(lib.SIGNATURES) <= declared


def test_defaults_match_python_mirror(L):
    p = ArtParams()
    assert L.art_params_default(C.byref(p)) == 0
    q = ArtParams.default()
    assert p.as_dict() == q.as_dict()
    assert C.sizeof(ArtParams) == 19 * 8 + 16 * 4


def test_max_radius_and_no_surface(L, oracle):
    from adiabatic_raytracer_b200 import lib
    p = ArtParams.default(theta_m=0.2)
    rc, r = lib.max_radius(p)
    assert rc == 0 and abs(r - oracle.max_radius(p)[1]) < 1e-12 * r
    rc, _ = lib.max_radius(ArtParams.default(theta_m=0.2, Mass_a=1e-4))
    assert rc == -3


def test_file_name_matches_julia_formatting(L):
    from adiabatic_raytracer_b200 import lib
    p = ArtParams.default(theta_m=0.2)
    name = lib.npy_name(p, 10001, "results", "7")
    assert name == ("results/npy/tree_MassAx_1.0e-5_AxionG_1.0e-12_ThetaM_0.2_rotPulsar_1.0_B0_1.0e14"
                    "_Ax_trajs_10001_N_Times_3_num_cutoff_5_MC_nodes_5_max_nodes_50_7.npy")
    p2 = ArtParams.default(theta_m=0.0, Mass_a=2.5e-6, omega_pul=6.283185307179586, B0=5e13, Ax_g=1e-9)
    assert "MassAx_2.5e-6_AxionG_1.0e-9_ThetaM_0.0_rotPulsar_6.283185307179586_B0_5.0e13" in lib.npy_name(p2, 5, "r", "")


def test_npy_writer_roundtrip_and_combine_files(L, tmp_path):
    """The files must keep working with the reference's Combine_Files.py logic (transpose, offset
    event ids, append) — re-stated here because /root/reference does not exist on the GPU box."""
    from adiabatic_raytracer_b200 import lib
    rng = np.random.default_rng(0)
    rows = rng.random((7, 13))
    rows[:, 0] = np.arange(1, 8)
    f1 = str(tmp_path / "tree_a.npy")
    f2 = str(tmp_path / "tree_b.npy")
    lib.write_npy(f1, rows)
    lib.write_npy(f2, rows * 2)
    a = np.load(f1)
    assert a.shape == (7, 13) and np.array_equal(a, rows) and np.isfortran(a)
    data = np.load(f1).T
    tmp = np.load(f2).T
    tmp[0, :] += data[0, -1]
    data = np.append(data, tmp, axis=1)
    assert data.shape == (13, 14) and data[0, 7] == rows[0, 0] * 2 + 7


def test_unsupported_switches_fail_loudly(L):
    """Parameter validation happens on the host (make_dev_params, reached through art_max_radius without a device):
    switches this build does not implement are refused with ART_ERR_ARG, the implemented ones are accepted."""
    from adiabatic_raytracer_b200 import lib
    for kw in (dict(bndry_lyr=0.5, melrose=0), dict(bndry_lyr=0.5, integrator=1), dict(n_maxSample=9), dict(integrator=7),
               dict(interp_points_ray=65), dict(cross_cap=0), dict(metric_interior_radius=12.0)):
        rc, _ = lib.max_radius(ArtParams.default(theta_m=0.2, **kw))
        assert rc == -1, kw
    for kw in (dict(melrose=0), dict(thick_surface=0), dict(isotropic=1), dict(flat=1), dict(integrator=1), dict(integrator=0),
               dict(bndry_lyr=0.5)):
        rc, maxR = lib.max_radius(ArtParams.default(theta_m=0.2, **kw))
        assert rc == 0 and maxR > 10.0, kw


def test_no_cpu_fallback(L):
    """Without a CUDA device art_create must fail with ART_ERR_CUDA (-2), never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from adiabatic_raytracer_b200 import lib
    with pytest.raises(lib.ArtError) as e:
        lib.Handle(ArtParams.default(theta_m=0.2), 0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_event_text_files_follow_the_reference_layout(L, tmp_path):
    """saveMode > 1: <dir_tag>/event/final_<tag> and event_<tag>, MainRunner.jl:593-610, 689-701, 733-735
    (numbers printed like Julia's string(): 1.0e-5, 0.25, integers without a decimal point)."""
    from adiabatic_raytracer_b200 import lib
    os.makedirs(tmp_path / "event")
    rows = np.zeros((2, 31))
    rows[0, [0, 1, 2, 3, 4, 5, 6, 13, 29, 30]] = [3, 1, 0.5, -1.25, 2.0, 1e-5, 299792.0, 2.5e-7, 1.0e-5, 0.0]
    rows[1, [0, 1, 2, 3, 4, 5, 6, 13, 29, 30]] = [3, 0, 1.5, 0.125, 0.25, 3.0, 1.0e7, 0.75, 9.5e-6, 1.5e-4]
    events = np.array([[3, 4.2e-4, 4.3e-4, 4.4e-4, 1.5e30, 10.0, -20.0, 30.5, 1e-6, -2e-6, 3e-6, 7]])
    lib.write_event_files(str(tmp_path), "tagX", rows, events)
    fin = open(tmp_path / "event" / "final_tagX").read().splitlines()
    assert fin[0] == "3 2.5e-7 1 0.5 -1.25 1.0e-5 2.0 1.0e-5 299792.0 0"
    assert fin[1] == "3 0.75 0 1.5 0.125 9.5e-6 0.25 3.0 1.0e7 0.00015"
    ev = open(tmp_path / "event" / "event_tagX").read().splitlines()
    assert ev == ["3 0.00042 0.00043 0.00044 1.5e30 10.0 -20.0 30.5 -1.0e-6 2.0e-6 -3.0e-6 10.0 -20.0 30.5 1.0e-6 -2.0e-6 3.0e-6 0.0 7"]
