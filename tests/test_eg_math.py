"""The hand-derived E_g Jacobian row (intrinsic3d_b200/csrc/i3d_math.cuh, compiled for the host by
tests/native/check_math.cpp) against the oracle's forward-mode Jets, on real rows of a scene."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OFFS = [(0, 0, 0), (0, 1, 0), (0, 2, 0), (0, 1, 1), (0, 0, 1), (0, 0, 2), (1, 0, 0), (1, 1, 0), (1, 0, 1), (2, 0, 0)]


def _harness():
    nat = os.path.join(ROOT, "tests", "native")
    so = os.path.join(nat, "libcheck_math.so")
    src = os.path.join(nat, "check_math.cpp")
    hdr = os.path.join(ROOT, "intrinsic3d_b200", "csrc", "i3d_math.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(so):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    return C.CDLL(so)


def _P(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _mine(L, coord, vs, ps, lum, sh, sdf, alb, pose, intr, dist):
    lum = np.ascontiguousarray(lum, np.float32)
    h, w = lum.shape
    coord = np.ascontiguousarray(coord, np.int32)
    arrs = [np.ascontiguousarray(a, np.float64) for a in (sh, sdf, alb, pose, intr, dist)]
    r = C.c_double()
    jd = np.zeros(29)
    jf = np.zeros(29, np.float32)
    L.i3dm_eval_eg(_P(coord, C.c_int32), C.c_double(vs), C.c_double(ps), C.c_int(w), C.c_int(h), _P(lum, C.c_float),
                   *[_P(a, C.c_double) for a in arrs], C.byref(r), _P(jd, C.c_double), _P(jf, C.c_float))
    return r.value, jd, jf


def _mine_voxel(L, coord, vs, ps, lum, sh, sdf, alb, pose, intr, dist, fn="i3dm_eval_eg_voxel"):
    """the voxel-owned evaluation path of the round-2 kernel k_eg_rows (frame-independent part hoisted, SO(3) right-Jacobian form)"""
    lum = np.ascontiguousarray(lum, np.float32)
    h, w = lum.shape
    coord = np.ascontiguousarray(coord, np.int32)
    arrs = [np.ascontiguousarray(a, np.float64) for a in (sh, sdf, alb, pose, intr, dist)]
    r = C.c_double()
    jf = np.zeros(29, np.float32)
    rc = getattr(L, fn)(_P(coord, C.c_int32), C.c_double(vs), C.c_double(ps), C.c_int(w), C.c_int(h), _P(lum, C.c_float),
                              *[_P(a, C.c_double) for a in arrs], C.byref(r), _P(jf, C.c_float))
    assert rc == 0, "cost-only instantiation disagrees with the build instantiation"
    return r.value, jf


@pytest.mark.parametrize("case", ["plain", "distorted", "small_angle", "tiny_angle", "pyr_scale", "border"])
def test_analytic_row_matches_jets(case, tiny_scene):
    from intrinsic3d_b200.ctypes_defs import default_params
    from oracle import Oracle, eval_eg
    L = _harness()
    s = tiny_scene
    o = Oracle(threads=4)
    o.load_scene(s)
    p = default_params()
    p.thres_shell = s["thres_shell"]
    p.build_only = 1
    o.gn_iteration(p)
    rows = o.rows(0)
    idx = {tuple(c): i for i, c in enumerate(s["xyz"])}
    rng = np.random.default_rng(3)
    dist = np.array([0.05, -0.02, 0.01, 0.003, -0.002]) if case == "distorted" else np.zeros(5)
    ps = 1.0
    lum_scale = None
    checked = 0
    for i in rng.choice(len(rows["voxel"]), 200, replace=False):
        v, f = rows["voxel"][i], rows["aux"][i]
        c = s["xyz"][v]
        sdf = np.array([s["sdf_refined"][idx[tuple(c + np.array(o_))]] for o_ in OFFS])
        alb = 0.6 + 0.1 * rng.standard_normal(4)
        pose = s["poses"][f].copy()
        intr = s["intr"].copy()
        lum = s["lum"][f]
        if case == "small_angle":
            pose[:3] = 1e-9 * rng.standard_normal(3)
        if case == "tiny_angle":       # general Rodrigues branch with a very small angle (series form of the SO(3) Jacobian)
            pose[:3] = 1e-4 * rng.standard_normal(3)
        if case == "border":           # crop the image so that many 4x4 neighbourhoods are clamped at the border
            intr = intr.copy(); intr[2] -= 40.0; intr[3] -= 25.0
        if case == "pyr_scale":
            ps = 0.5
            intr = intr * 2.0       # full-resolution intrinsics, level-1 image
        vs = float(s["voxel_size"])
        r0, j0 = eval_eg(c, vs, ps, lum, s["sh"][v], sdf, alb, pose, intr, dist)
        r1, jd, jf = _mine(L, c, vs, ps, lum, s["sh"][v], sdf, alb, pose, intr, dist)
        if r0 == 0.0:
            assert r1 == 0.0
            continue
        checked += 1
        assert abs(r1 - r0) <= 1e-12 * abs(r0)
        sc = np.abs(j0).max()
        assert np.abs(jd - j0).max() <= 1e-6 * sc       # f64 chain rule (image gradient passed as float)
        assert np.abs(jf - j0).max() <= 2e-5 * sc       # f32 derivative pass (round-1 formulation)
        r2, jv = _mine_voxel(L, c, vs, ps, lum, s["sh"][v], sdf, alb, pose, intr, dist)
        assert abs(r2 - r0) <= 1e-11 * abs(r0)          # voxel-owned evaluation (thread per voxel)
        assert np.abs(jv - j0).max() <= 2e-5 * sc
        # the same row with the per-voxel state read through the shared-memory views of k_eg_rows: bit-identical
        r3, jw = _mine_voxel(L, c, vs, ps, lum, s["sh"][v], sdf, alb, pose, intr, dist, fn="i3dm_eval_eg_voxel_views")
        assert r3 == r2 and np.array_equal(jw, jv)
    assert checked > 20
