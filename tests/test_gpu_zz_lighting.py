"""GPU parity for the SVSH lighting path (i3d_estimate_lighting = LightingSVSH::estimate + computeVoxelShCoeffs) against
the float64 CPU oracle.  Everything is float64 on both sides; the engine solves the reduced per-subvolume normal
equations, the oracle the explicit row problem, so:
  subvolume set / numbering, row and pair counts, has_sh mask : exact
  weight sum, initial cost                                    : rel <= 1e-11
  LM / CG iteration counts                                    : equal
  subvolume SH, per-voxel SH                                  : <= 1e-8 of max |sh|
and the E_g residuals built from the estimated per-voxel SH match the oracle's (rel <= 1e-8)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lighting_params(mod, scene, **kw):
    P = mod.default_lighting_params()
    P.thres_shell = float(scene["thres_shell"])
    P.subvolume_size = 0.03
    for k, v in kw.items():
        setattr(P, k, v)
    return P


def _both(scene, **kw):
    import oracle
    from intrinsic3d_b200 import engine
    e = engine.Engine(0)
    e.load_scene(scene)
    o = oracle.Oracle(threads=8)
    o.load_scene(scene)
    ie = e.estimate_lighting(_lighting_params(engine, scene, **kw))
    io = o.estimate_lighting(_lighting_params(oracle, scene, **kw))
    return e, o, ie, io


def _compare(e, o, ie, io, tol=1e-8):
    assert ie.usable == io.usable == 1
    idx_e, sh_e = e.download_lighting()
    idx_o, sh_o = o.lighting()
    assert np.array_equal(idx_e, idx_o)
    assert ie.num_subvolumes == io.num_subvolumes == len(idx_o)
    assert ie.num_data_rows == io.num_data_rows and ie.num_reg_pairs == io.num_reg_pairs
    np.testing.assert_allclose(ie.sum_data_weights, io.sum_data_weights, rtol=1e-11)
    np.testing.assert_allclose(ie.cost_initial, io.cost_initial, rtol=1e-11)
    assert (ie.lm_iterations, ie.num_successful_steps, ie.cg_iterations_total, ie.termination) == \
           (io.lm_iterations, io.num_successful_steps, io.cg_iterations_total, io.termination)
    np.testing.assert_allclose(ie.cost_final, io.cost_final, rtol=1e-9)
    np.testing.assert_allclose(ie.trust_region_radius, io.trust_region_radius, rtol=1e-9)
    ref = np.abs(sh_o).max()
    assert np.abs(sh_e - sh_o).max() <= tol * ref
    vsh_e, has_e = e.download_voxel_sh()
    vsh_o, has_o = o.voxel_sh()
    assert np.array_equal(has_e, has_o) and has_o.sum() > 100
    assert np.abs(vsh_e - vsh_o).max() <= tol * ref
    return ref


@pytest.mark.parametrize("weighted,lambda_reg,size", [(1, 10.0, 0.03), (0, 0.5, 0.05), (1, 10.0, 0.2)])
def test_lighting_matches_oracle(weighted, lambda_reg, size, small_scene):
    e, o, ie, io = _both(small_scene, weighted=weighted, lambda_reg=lambda_reg, subvolume_size=size)
    _compare(e, o, ie, io)


def test_estimated_sh_feeds_the_gn_iteration(tiny_scene):
    """i3d_estimate_lighting leaves the per-voxel blend as the `sh` input of i3d_gn_iteration (replacing i3d_set_sh)."""
    from intrinsic3d_b200.ctypes_defs import default_params
    e, o, ie, io = _both(tiny_scene, subvolume_size=0.02)
    _compare(e, o, ie, io)
    p = default_params()
    p.thres_shell = tiny_scene["thres_shell"]
    p.build_only = 1
    e.gn_iteration(p)
    o.gn_iteration(p)
    rows = e.debug_rows(want_jac=False)
    eg = o.rows(0)
    me = {(int(v), int(f)): i for i, (v, f) in enumerate(zip(rows["voxel"], rows["frame"])) if f >= 0}
    mo = {(int(v), int(f)): i for i, (v, f) in enumerate(zip(eg["voxel"], eg["aux"]))}
    assert set(me) == set(mo) and len(mo) > 500
    ie_ = np.array([me[k] for k in mo]); io_ = np.array([mo[k] for k in mo])
    assert np.max(np.abs(rows["residual"][ie_] - eg["residual"][io_]) / np.abs(eg["residual"][io_])) < 1e-8
    # and it differs from what the scene's own SH would have produced (the estimate really replaced it)
    vsh, _ = e.download_voxel_sh()
    assert np.abs(vsh - tiny_scene["sh"]).max() > 1e-3


def test_lighting_edge_cases(tiny_scene):
    import oracle
    from intrinsic3d_b200 import engine
    s = tiny_scene
    e = engine.Engine(0)
    e.upload_grid(s["xyz"], s["sdf0"], s["sdf_refined"], s["albedo"], s["weight"], s["rgb"], s["voxel_size"])
    # thres_shell <= 0: estimate() returns false, nothing computed
    info = e.estimate_lighting(_lighting_params(engine, s, thres_shell=0.0))
    assert info.usable == 0 and info.num_subvolumes == 0
    with pytest.raises(RuntimeError):
        e.download_voxel_sh()
    # non-positive subvolume size is rejected loudly
    with pytest.raises(RuntimeError):
        e.estimate_lighting(_lighting_params(engine, s, subvolume_size=0.0))
    # iteration cap: NO_CONVERGENCE is still usable
    info = e.estimate_lighting(_lighting_params(engine, s, max_iterations=1))
    assert info.usable == 1 and info.termination == 1 and info.lm_iterations == 1
    # empty data term (no voxel inside the shell): zero lighting, convergence at iteration 0
    info = e.estimate_lighting(_lighting_params(engine, s, thres_shell=1e-9))
    assert info.usable == 1 and info.num_data_rows == 0 and info.lm_iterations == 0
    idx, sh = e.download_lighting()
    assert np.all(sh == 0.0)
    # voxels with weight 0 and black / zero-albedo voxels drop out exactly as in the oracle
    s2 = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in s.items()}
    rng = np.random.default_rng(0)
    s2["weight"][rng.choice(len(s2["weight"]), 200, replace=False)] = 0.0
    s2["albedo"][rng.choice(len(s2["albedo"]), 200, replace=False)] = 0.0
    s2["albedo"][rng.choice(len(s2["albedo"]), 50, replace=False)] = np.nan
    e2, o2, ie, io = _both(s2, subvolume_size=0.02)
    _compare(e2, o2, ie, io)


def test_engine_matches_golden_lighting():
    """Committed fixture tests/golden/tiny_lighting.npz (oracle output on the grid of tiny_gn.npz)."""
    from intrinsic3d_b200 import engine
    from test_golden import _check_lighting, _lighting_params as golden_params, _load_lighting
    L, s = _load_lighting()
    e = engine.Engine(0)
    e.upload_grid(s["xyz"], s["sdf0"], s["sdf_refined"], s["albedo"], s["weight"], s["rgb"], s["voxel_size"])
    info = e.estimate_lighting(golden_params(engine, L, s))
    idx, sh = e.download_lighting()
    vsh, has = e.download_voxel_sh()
    _check_lighting(L, info, idx, sh, vsh, has, 1e-8)


def test_cpp_lighting_svsh_matches_engine(small_scene):
    """nv::LightingSVSH (reference-shaped C++ API over the C-ABI): estimate(), shCoeffs(), subvolumes(), computeVoxelShCoeffs()
    reproduce the direct C-ABI call, and its host-side interpolate() reproduces the device blend."""
    import ctypes as C
    import os
    from intrinsic3d_b200 import engine
    s = small_scene
    e = engine.Engine(0)
    e.upload_grid(s["xyz"], s["sdf0"], s["sdf_refined"], s["albedo"], s["weight"], s["rgb"], s["voxel_size"])
    P = _lighting_params(engine, s, subvolume_size=0.04, lambda_reg=5.0)
    info = e.estimate_lighting(P)
    idx, sh = e.download_lighting()
    vsh, has = e.download_voxel_sh()

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    H = C.CDLL(os.path.join(root, "intrinsic3d_b200", "libi3d_host.so"))
    n = s["xyz"].shape[0]

    def ptr(a, t):
        return a.ctypes.data_as(C.POINTER(t))
    xyz = np.ascontiguousarray(s["xyz"], np.int32)
    arrs = [np.ascontiguousarray(s[k], np.float64) for k in ("sdf0", "sdf_refined", "albedo")]
    wgt = np.ascontiguousarray(s["weight"], np.float32)
    rgb = np.ascontiguousarray(s["rgb"], np.uint8)
    cap = 4096
    nsub = np.zeros(1, np.int64)
    idx2 = np.zeros((cap, 3), np.int32)
    sh2 = np.zeros((cap, 9), np.float64)
    vsh2 = np.zeros((n, 9), np.float64)
    has2 = np.zeros(n, np.uint8)
    err = np.zeros(1, np.float64)
    rc = H.i3dh_run_lighting(C.c_int64(n), ptr(xyz, C.c_int32), ptr(arrs[0], C.c_double), ptr(arrs[1], C.c_double), ptr(arrs[2], C.c_double),
                             ptr(wgt, C.c_float), ptr(rgb, C.c_uint8), C.c_float(float(s["voxel_size"])), C.c_float(P.subvolume_size),
                             C.c_double(P.lambda_reg), C.c_double(P.thres_shell), C.c_int32(1), C.c_int64(cap), ptr(nsub, C.c_int64),
                             ptr(idx2, C.c_int32), ptr(sh2, C.c_double), ptr(vsh2, C.c_double), ptr(has2, C.c_uint8), ptr(err, C.c_double))
    assert rc == 0
    S = int(nsub[0])
    assert S == info.num_subvolumes and np.array_equal(idx2[:S], idx)
    ref = np.abs(sh).max()
    assert np.abs(sh2[:S] - sh).max() <= 1e-9 * ref          # two device runs differ only by the order of the double atomics
    assert np.array_equal(has2, has)
    assert np.abs(vsh2 - vsh).max() <= 1e-9 * ref
    assert err[0] <= 1e-12 * ref                              # host interpolate() == device blend
