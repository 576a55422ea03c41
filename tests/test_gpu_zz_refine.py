"""nv::Intrinsic3D::refine (reference-shaped orchestrator on ONE resident engine) against the same schedule driven from Python through
the C-ABI: convert -> initial recolouring -> per grid level {thin-shell pruning -> per pyramid level {lighting, GN iterations,
recolouring} -> upsample}.  Two runs of the engine differ in the last float bits (atomic accumulation order), and a voxel on the
pruning threshold can flip, so the comparison is on counts (within 0.5 %) and, for the voxels both runs hold, statistical."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lam(it, n, a, b):
    return a if n <= 1 else a + (b - a) * it / (n - 1)          # computeVaryingLambda (include/nv/refinement/cost.h)


def test_cpp_refine_matches_python_driven_schedule():
    from intrinsic3d_b200 import engine
    from intrinsic3d_b200.ctypes_defs import default_params
    from intrinsic3d_b200.scene import make_color_frames, make_scene
    s = make_scene(radius_vox=12.0, frames=6, width=160, height=120, voxel_size=0.008, band=3.0, seed=6)
    col = make_color_frames(s)
    F, H, W = s["lum"].shape
    lum1 = s["lum"].reshape(F, H // 2, 2, W // 2, 2).mean((2, 4)).astype(np.float32)
    dep1 = np.ascontiguousarray(s["depth"][:, ::2, ::2])
    GL, RL, ITS = 2, 2, 2
    factor0, factor1 = 2.0, 1.0
    lam = dict(g=0.2, r0=80.0, r1=10.0, s0=120.0, s1=10.0, a=0.1)
    sub_size, sh_reg, occl, K = 0.06, 10.0, 0.02, 5
    intr0 = np.ascontiguousarray(s["intr"], np.float64)
    keep = s["weight"] > 0

    # ---------------- Python-driven schedule on one engine
    e = engine.Engine(0)
    e.upload_grid(s["xyz"][keep], s["sdf0"][keep].astype(np.float32).astype(np.float64), s["sdf0"][keep].astype(np.float32).astype(np.float64),
                  np.full(int(keep.sum()), 0.6), s["weight"][keep], s["rgb"][keep], s["voxel_size"])
    e.upload_frames(s["lum"], s["depth"], 1.0)
    e.upload_color_frames(col)
    e.set_camera(s["poses"], intr0, np.zeros(5))
    e.recompute_colors(occl, K)
    vs = float(np.float32(s["voxel_size"]))
    LP = engine.default_lighting_params()
    LP.subvolume_size = sub_size; LP.lambda_reg = sh_reg; LP.weighted = 1
    level = 0
    for gl in range(GL - 1, -1, -1):
        fac = _lam(GL - 1 - gl, GL, factor0, factor1)
        thres = fac * vs
        e.clear_voxels_outside_thin_shell(thres)
        for rl in range(RL - 1, -1, -1):
            if rl > 0 and gl < GL - 1:
                continue
            if rl != level:
                e.upload_frames(lum1 if rl == 1 else s["lum"], dep1 if rl == 1 else s["depth"], 1.0 / 2 ** rl)
                level = rl
            LP.thres_shell = thres
            li = e.estimate_lighting(LP)
            assert li.usable == 1
            for it in range(ITS):
                p = default_params()
                p.thres_shell = thres; p.occlusion_distance = occl; p.num_observations = K
                p.lambda_[0] = lam["g"]; p.lambda_[1] = _lam(it, ITS, lam["r0"], lam["r1"]); p.lambda_[2] = _lam(it, ITS, lam["s0"], lam["s1"]); p.lambda_[3] = lam["a"]
                e.gn_iteration(p)
            if level != 0:
                e.upload_frames(s["lum"], s["depth"], 1.0); level = 0
                e.upload_color_frames(col)
            e.recompute_colors(occl, K)
        if gl > 0:
            e.upsample_grid()
            vs = float(np.float32(np.float32(vs) * np.float32(0.5)))
    ref = e.download_grid()
    ref_state = e.download_state()

    # ---------------- C++ orchestrator
    Hh = C.CDLL(os.path.join(ROOT, "intrinsic3d_b200", "libi3d_host.so"))

    def ptr(a, t):
        return a.ctypes.data_as(C.POINTER(t))
    n = len(s["xyz"])
    xyz = np.ascontiguousarray(s["xyz"], np.int32)
    sdf = np.ascontiguousarray(s["sdf0"], np.float32)
    wgt = np.ascontiguousarray(s["weight"], np.float32)
    rgb = np.ascontiguousarray(s["rgb"], np.uint8)
    lum0 = np.ascontiguousarray(s["lum"], np.float32); dep0 = np.ascontiguousarray(s["depth"], np.float32)
    Wl = np.array([W, W // 2], np.int32); Hl = np.array([H, H // 2], np.int32)
    lum_ptrs = (C.POINTER(C.c_float) * 2)(ptr(lum0, C.c_float), ptr(lum1, C.c_float))
    dep_ptrs = (C.POINTER(C.c_float) * 2)(ptr(dep0, C.c_float), ptr(dep1, C.c_float))
    colc = np.ascontiguousarray(col, np.uint8)
    poses = np.ascontiguousarray(s["poses"], np.float64).copy()
    intr = intr0.copy(); dist = np.zeros(5)
    cfg = np.array([GL, RL, factor0, factor1, 1, occl, K, sub_size, sh_reg, ITS, 50, lam["g"], lam["r0"], lam["r1"], lam["s0"], lam["s1"], lam["a"]], np.float64)
    cap = 8 * n
    out = dict(xyz=np.zeros((cap, 3), np.int32), sdf0=np.zeros(cap), sdf=np.zeros(cap), alb=np.zeros(cap), w=np.zeros(cap, np.float32), rgb=np.zeros((cap, 3), np.uint8))
    m = C.c_int64(0); vso = C.c_float(0); calls = C.c_int32(0)
    rc = Hh.i3dh_run_refine(C.c_int64(n), ptr(xyz, C.c_int32), ptr(sdf, C.c_float), ptr(wgt, C.c_float), ptr(rgb, C.c_uint8), C.c_float(float(s["voxel_size"])),
                            C.c_int32(F), C.c_int32(2), ptr(Wl, C.c_int32), ptr(Hl, C.c_int32), lum_ptrs, dep_ptrs, ptr(colc, C.c_uint8), ptr(poses, C.c_double),
                            ptr(intr, C.c_double), ptr(dist, C.c_double), ptr(cfg, C.c_double), C.c_int64(cap), C.byref(m), ptr(out["xyz"], C.c_int32),
                            ptr(out["sdf0"], C.c_double), ptr(out["sdf"], C.c_double), ptr(out["alb"], C.c_double), ptr(out["w"], C.c_float), ptr(out["rgb"], C.c_uint8),
                            C.byref(vso), C.byref(calls))
    assert rc == 0
    M = int(m.value)
    assert calls.value == 3                                # (gl 1: rl 1, rl 0) + (gl 0: rl 0)
    assert vso.value == ref["voxel_size"] == np.float32(np.float32(s["voxel_size"]) * np.float32(0.5))
    assert abs(M - len(ref["xyz"])) <= 0.005 * len(ref["xyz"]) and M > 8 * 0.1 * n
    a = {tuple(c): i for i, c in enumerate(ref["xyz"])}
    common = [(a[tuple(c)], i) for i, c in enumerate(out["xyz"][:M]) if tuple(c) in a]
    assert len(common) >= 0.995 * M
    ia, ib = np.array(common).T
    moved = np.abs(ref["sdf_refined"] - ref["sdf0"]).max()
    d = np.abs(ref["sdf_refined"][ia] - out["sdf"][:M][ib])
    assert moved > 0 and (d <= 1e-3 * moved).mean() > 0.98, ((d <= 1e-3 * moved).mean(), d.max(), moved)
    da = np.abs(ref["albedo"][ia] - out["alb"][:M][ib])
    assert (da <= 1e-3).mean() > 0.98
    dc = np.abs(ref["rgb"][ia].astype(int) - out["rgb"][:M][ib].astype(int)).max(1)
    assert (dc <= 1).mean() > 0.97
    assert np.abs(poses - ref_state["poses"]).max() < 1e-4 and np.abs(intr - ref_state["intr"]).max() < 1e-2
    assert not np.allclose(poses, s["poses"])
