"""Committed golden vectors (tests/golden/tiny_gn.npz, made by tests/golden/make_golden.py from the CPU oracle):
 - CPU: the oracle still reproduces them (regression pin; the reference itself has no fixtures),
 - GPU: the CUDA engine, through the C-ABI, matches them to the parity tolerances."""
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    g = np.load(os.path.join(ROOT, "tests", "golden", "tiny_gn.npz"))
    s = {k: g[k] for k in ("xyz", "sdf0", "sdf_refined", "albedo", "weight", "rgb", "lum", "depth", "poses", "intr", "dist", "sh")}
    s["voxel_size"] = np.float32(g["voxel_size"])
    s["thres_shell"] = float(g["thres_shell"])
    return g, s


def _params(g, s):
    from intrinsic3d_b200.ctypes_defs import default_params
    p = default_params()
    p.thres_shell = s["thres_shell"]
    p.forced_cg_iterations = int(g["forced_cg"])
    return p


def test_oracle_reproduces_golden():
    from oracle import Oracle
    g, s = _load()
    o = Oracle(threads=4)
    o.load_scene(s)
    info = o.gn_iteration(_params(g, s))
    fr, w, act = o.observations(5)
    assert np.array_equal(act, g["active"]) and np.array_equal(fr, g["obs_frames"])
    assert np.array_equal(w.view(np.uint32), g["obs_weights"].view(np.uint32))
    eg = o.rows(0)
    assert np.array_equal(eg["voxel"], g["eg_voxel"]) and np.array_equal(eg["aux"], g["eg_frame"])
    np.testing.assert_allclose(eg["residual"], g["eg_residual"], rtol=1e-10)
    assert list(info.type_residuals) == list(g["type_residuals"])
    np.testing.assert_allclose(info.cost_initial, float(g["cost_initial"]), rtol=1e-10)
    np.testing.assert_allclose(info.cost_final, float(g["cost_final"]), rtol=1e-8)
    st = o.state()
    ref = np.abs(g["step"]).max()
    assert np.abs(st["sdf_refined"] - g["out_sdf"]).max() <= 1e-7 * ref


@pytest.mark.gpu
def test_engine_matches_golden():
    from intrinsic3d_b200.engine import Engine
    g, s = _load()
    e = Engine(0)
    e.load_scene(s)
    info = e.gn_iteration(_params(g, s))
    fr, w, act = e.debug_observations(5)
    assert np.array_equal(act, g["active"]) and np.array_equal(fr, g["obs_frames"])
    assert np.array_equal(w.view(np.uint32), g["obs_weights"].view(np.uint32))
    rows = e.debug_rows()
    me = {(int(v), int(f)): i for i, (v, f) in enumerate(zip(rows["voxel"], rows["frame"])) if f >= 0}
    mo = {(int(v), int(f)): i for i, (v, f) in enumerate(zip(g["eg_voxel"], g["eg_frame"]))}
    assert set(me) == set(mo)
    ie = np.array([me[k] for k in mo]); io = np.array([mo[k] for k in mo])
    assert np.max(np.abs(rows["residual"][ie] - g["eg_residual"][io]) / np.abs(g["eg_residual"][io])) < 1e-9
    Je, Jo = rows["J"][:, ie].T.astype(np.float64), g["eg_jacobian"][io].astype(np.float64)
    assert np.max(np.abs(Je - Jo) / np.abs(Jo).max(axis=1, keepdims=True)) < 1e-4
    assert list(info.type_residuals) == list(g["type_residuals"])
    np.testing.assert_allclose(list(info.type_sum_weights), g["type_sum_weights"], rtol=1e-9)
    np.testing.assert_allclose(info.cost_initial, float(g["cost_initial"]), rtol=1e-9)
    assert info.step_accepted == int(g["accepted"]) and info.lm_iterations == int(g["lm_iterations"])
    np.testing.assert_allclose(info.model_cost_change[0], float(g["model_cost_change"]), rtol=5e-4)
    np.testing.assert_allclose(info.cost_final, float(g["cost_final"]), rtol=5e-4)
    st = e.download_state()
    n = s["xyz"].shape[0]
    step = g["step"].astype(np.float64)
    assert np.abs(st["sdf_refined"] - g["out_sdf"]).max() <= 1e-3 * np.abs(step[:n]).max()
    assert np.abs(st["albedo"] - g["out_albedo"]).max() <= 1e-3 * np.abs(step[n:2 * n]).max()
    assert np.abs(st["poses"] - g["out_poses"]).max() <= 1e-3 * np.abs(step[2 * n:2 * n + 30]).max()


# ---- SVSH lighting (tests/golden/tiny_lighting.npz: oracle output on the grid of tiny_gn.npz) ----
def _load_lighting():
    g, s = _load()
    L = np.load(os.path.join(ROOT, "tests", "golden", "tiny_lighting.npz"))
    return L, s


def _lighting_params(mod, L, s):
    P = mod.default_lighting_params()
    P.thres_shell = s["thres_shell"]
    P.subvolume_size = float(L["subvolume_size"])
    P.lambda_reg = float(L["lambda_reg"])
    return P


def _check_lighting(L, info, idx, sh, vsh, has, tol):
    assert info.usable == 1 and info.termination == int(L["termination"])
    assert np.array_equal(idx, L["sub_index"])
    assert info.num_data_rows == int(L["num_data_rows"]) and info.num_reg_pairs == int(L["num_reg_pairs"])
    np.testing.assert_allclose(info.sum_data_weights, float(L["sum_data_weights"]), rtol=1e-12)
    np.testing.assert_allclose(info.cost_initial, float(L["cost_initial"]), rtol=1e-11)
    assert info.lm_iterations == int(L["lm_iterations"]) and info.num_successful_steps == int(L["num_successful_steps"])
    assert info.cg_iterations_total == int(L["cg_iterations_total"])
    np.testing.assert_allclose(info.cost_final, float(L["cost_final"]), rtol=1e-9)
    ref = np.abs(L["sub_sh"]).max()
    assert np.abs(sh - L["sub_sh"]).max() <= tol * ref
    assert np.array_equal(has, L["has_sh"])
    assert np.abs(vsh - L["voxel_sh"]).max() <= tol * ref


def test_oracle_reproduces_golden_lighting():
    import oracle
    L, s = _load_lighting()
    o = oracle.Oracle(threads=2)
    o.set_grid(s)
    info = o.estimate_lighting(_lighting_params(oracle, L, s))
    idx, sh = o.lighting()
    vsh, has = o.voxel_sh()
    _check_lighting(L, info, idx, sh, vsh, has, 1e-12)


# ---- voxel recolouring (tests/golden/tiny_recolor.npz: oracle output for the scene of tiny_gn.npz) ----
def _load_recolor():
    g, s = _load()
    return np.load(os.path.join(ROOT, "tests", "golden", "tiny_recolor.npz")), s


def test_oracle_reproduces_golden_recolor():
    import oracle
    R, s = _load_recolor()
    for tag, K in (("k", int(R["K"])), ("all", 0)):
        o = oracle.Oracle(threads=2)
        o.load_scene(s)
        o.set_color_frames(R["color"])
        cnt = o.recompute_colors(float(R["occlusion"]), K)
        assert list(cnt) == list(R["counts_" + tag])
        assert np.array_equal(o.colors(), R["rgb_" + tag])


# ---- grid-level transitions (tests/golden/tiny_gridops.npz: oracle output for the grid of tiny_gn.npz) ----
def _load_gridops():
    g, s = _load()
    G = np.load(os.path.join(ROOT, "tests", "golden", "tiny_gridops.npz"))
    s = dict(s)
    s["weight"] = G["weight_in"]
    return G, s


def _check_gridops(G, obj, grid_of):
    """obj: an Oracle or an Engine with the golden scene loaded; grid_of(obj) -> dict of arrays."""
    vs = float(G["shell"])
    assert obj.clear_voxels_outside_thin_shell(vs) == int(G["counts"][0])
    g1 = grid_of(obj)
    assert np.array_equal(g1["xyz"], G["prune_xyz"]) and np.array_equal(g1["sdf_refined"], G["prune_sdf"])
    assert obj.upsample_grid() == int(G["counts"][1])
    g2 = grid_of(obj)
    assert g2["voxel_size"] == G["up_voxel_size"]
    assert np.array_equal(g2["xyz"][:4096], G["up_xyz_head"])
    assert np.array_equal(g2["xyz"][::8], 2 * G["prune_xyz"])                       # child (0,0,0) of every parent, in parent order
    for key, name in (("sdf_refined", "up_sdf"), ("sdf0", "up_sdf0"), ("albedo", "up_albedo")):
        assert np.array_equal(g2[key], G[name].astype(np.float64)), key
    assert np.array_equal(g2["weight"], G["up_weight"]) and np.array_equal(g2["rgb"], G["up_rgb"])
    assert obj.clear_voxels_outside_thin_shell(0.5 * vs) == int(G["counts"][2])
    g3 = grid_of(obj)
    assert np.array_equal(g3["xyz"], G["final_xyz"]) and np.array_equal(g3["rgb"], G["final_rgb"])


def test_oracle_reproduces_golden_gridops():
    import oracle
    G, s = _load_gridops()
    o = oracle.Oracle(threads=2)
    o.load_scene(s)
    _check_gridops(G, o, lambda x: x.grid())
