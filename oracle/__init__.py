"""ctypes wrapper around oracle/liboracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from intrinsic3d_b200.ctypes_defs import I3DIterInfo, I3DLightingInfo, I3DLightingParams, I3DParams

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.i3do_create.restype = C.c_void_p
        L.i3do_last_error.restype = C.c_char_p
        L.i3do_num_rows.restype = C.c_int64
        L.i3do_num_subvolumes.restype = C.c_int64
        L.i3do_num_voxels.restype = C.c_int64
        L.i3do_num_lighting_rows.restype = C.c_int64
        for name in ("i3do_destroy", "i3do_last_error", "i3do_set_threads", "i3do_set_grid", "i3do_set_frames",
                     "i3do_set_camera", "i3do_set_sh", "i3do_gn_iteration", "i3do_get_state", "i3do_num_rows",
                     "i3do_get_rows", "i3do_get_eg_jacobian", "i3do_get_observations", "i3do_get_step"):
            getattr(L, name).argtypes = None
        _LIB = L
    return _LIB


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class Oracle:
    """Float64 CPU restatement of Optimizer::optimize's outer iteration (see oracle.cpp header)."""

    def __init__(self, threads: int = 8, parallel_cg: bool = False):
        self.L = lib()
        self.h = C.c_void_p(self.L.i3do_create())
        self.L.i3do_set_threads(self.h, C.c_int(threads), C.c_int(int(parallel_cg)))
        self.n = 0
        self.F = 0

    def __del__(self):
        if getattr(self, "h", None):
            self.L.i3do_destroy(self.h)
            self.h = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.i3do_last_error(self.h).decode())

    def load_scene(self, s):
        """s: dict from intrinsic3d_b200.scene.make_scene"""
        n = int(s["xyz"].shape[0])
        self.n = n
        self._keep = [np.ascontiguousarray(s["xyz"], np.int32), np.ascontiguousarray(s["sdf0"], np.float64),
                      np.ascontiguousarray(s["sdf_refined"], np.float64), np.ascontiguousarray(s["albedo"], np.float64),
                      np.ascontiguousarray(s["weight"], np.float32), np.ascontiguousarray(s["rgb"], np.uint8)]
        k = self._keep
        self._check(self.L.i3do_set_grid(self.h, C.c_int64(n), _p(k[0], C.c_int32), _p(k[1], C.c_double), _p(k[2], C.c_double),
                                         _p(k[3], C.c_double), _p(k[4], C.c_float), _p(k[5], C.c_uint8),
                                         C.c_float(float(s["voxel_size"]))))
        lum = np.ascontiguousarray(s["lum"], np.float32)
        depth = np.ascontiguousarray(s["depth"], np.float32)
        F, H, W = lum.shape
        self.F = F
        self._check(self.L.i3do_set_frames(self.h, C.c_int(F), C.c_int(W), C.c_int(H), _p(lum, C.c_float), _p(depth, C.c_float),
                                           C.c_double(float(s.get("pyr_scale", 1.0)))))
        self.set_camera(s["poses"], s["intr"], s["dist"])
        sh = np.ascontiguousarray(s["sh"], np.float64)
        self._check(self.L.i3do_set_sh(self.h, _p(sh, C.c_double)))

    def set_frames(self, lum, depth, pyr_scale=1.0):
        """Frames of another pyramid level (prepareRgbdLevel): same frame count, new size / scale."""
        lum = np.ascontiguousarray(lum, np.float32)
        depth = np.ascontiguousarray(depth, np.float32)
        F, H, W = lum.shape
        self.F = F
        self._frames = (lum, depth)
        self._check(self.L.i3do_set_frames(self.h, C.c_int(F), C.c_int(W), C.c_int(H), _p(lum, C.c_float), _p(depth, C.c_float), C.c_double(float(pyr_scale))))

    def set_camera(self, poses, intr, dist):
        poses = np.ascontiguousarray(poses, np.float64)
        intr = np.ascontiguousarray(intr, np.float64)
        dist = np.ascontiguousarray(dist, np.float64)
        self._check(self.L.i3do_set_camera(self.h, _p(poses, C.c_double), _p(intr, C.c_double), _p(dist, C.c_double)))

    def gn_iteration(self, params: I3DParams) -> I3DIterInfo:
        info = I3DIterInfo()
        self._check(self.L.i3do_gn_iteration(self.h, C.byref(params), C.byref(info)))
        return info

    def state(self):
        sdf = np.empty(self.n, np.float64)
        alb = np.empty(self.n, np.float64)
        poses = np.empty((self.F, 6), np.float64)
        intr = np.empty(4, np.float64)
        dist = np.empty(5, np.float64)
        self.L.i3do_get_state(self.h, _p(sdf, C.c_double), _p(alb, C.c_double), _p(poses, C.c_double), _p(intr, C.c_double),
                              _p(dist, C.c_double))
        return dict(sdf_refined=sdf, albedo=alb, poses=poses, intr=intr, dist=dist)

    def rows(self, type_id: int):
        m = int(self.L.i3do_num_rows(self.h, C.c_int(type_id)))
        voxel = np.empty(m, np.int32)
        aux = np.empty(m, np.int32)
        res = np.empty(m, np.float64)
        w = np.empty(m, np.float64)
        wr = np.empty(m, np.float64)
        self.L.i3do_get_rows(self.h, C.c_int(type_id), _p(voxel, C.c_int32), _p(aux, C.c_int32), _p(res, C.c_double),
                             _p(w, C.c_double), _p(wr, C.c_double))
        return dict(voxel=voxel, aux=aux, residual=res, weight=w, raw_weight=wr)

    def eg_jacobian(self):
        m = int(self.L.i3do_num_rows(self.h, C.c_int(0)))
        J = np.empty((m, 29), np.float64)
        self.L.i3do_get_eg_jacobian(self.h, _p(J, C.c_double))
        return J

    def observations(self, K: int):
        fr = np.empty((self.n, K), np.int32)
        w = np.empty((self.n, K), np.float32)
        act = np.empty(self.n, np.uint8)
        self._check(self.L.i3do_get_observations(self.h, C.c_int(K), _p(fr, C.c_int32), _p(w, C.c_float), _p(act, C.c_uint8)))
        return fr, w, act

    # ---- grid-level transitions (SDFAlgorithms::clearVoxelsOutsideThinShell / upsample) ----
    def clear_voxels_outside_thin_shell(self, thres_shell: float) -> int:
        self._check(self.L.i3do_clear_voxels_outside_thin_shell(self.h, C.c_double(thres_shell)))
        self.n = int(self.L.i3do_num_voxels(self.h))
        return self.n

    def upsample_grid(self) -> int:
        self._check(self.L.i3do_upsample_grid(self.h))
        self.n = int(self.L.i3do_num_voxels(self.h))
        return self.n

    def grid(self):
        n = int(self.L.i3do_num_voxels(self.h))
        out = dict(xyz=np.empty((n, 3), np.int32), sdf0=np.empty(n, np.float64), sdf_refined=np.empty(n, np.float64), albedo=np.empty(n, np.float64),
                   weight=np.empty(n, np.float32), rgb=np.empty((n, 3), np.uint8))
        vs = C.c_float(0)
        self.L.i3do_get_grid(self.h, _p(out["xyz"], C.c_int32), _p(out["sdf0"], C.c_double), _p(out["sdf_refined"], C.c_double), _p(out["albedo"], C.c_double),
                             _p(out["weight"], C.c_float), _p(out["rgb"], C.c_uint8), C.byref(vs))
        out["voxel_size"] = np.float32(vs.value)
        return out

    # ---- voxel recolouring (Intrinsic3D::recomputeColors) ----
    def set_color_frames(self, bgr):
        self._color = np.ascontiguousarray(bgr, np.uint8)
        self._check(self.L.i3do_set_color_frames(self.h, _p(self._color, C.c_uint8)))

    def recompute_colors(self, max_occlusion_distance: float = 0.02, max_num_observations: int = 5):
        cnt = np.zeros(2, np.int64)
        self._check(self.L.i3do_recompute_colors(self.h, C.c_float(max_occlusion_distance), C.c_int(max_num_observations), _p(cnt, C.c_int64)))
        return int(cnt[0]), int(cnt[1])

    def colors(self):
        rgb = np.empty((self.n, 3), np.uint8)
        self.L.i3do_get_colors(self.h, _p(rgb, C.c_uint8))
        return rgb

    # ---- SVSH lighting (LightingSVSH::estimate + computeVoxelShCoeffs) ----
    def set_grid(self, s):
        """Grid only (the lighting estimate needs neither frames nor camera)."""
        n = int(s["xyz"].shape[0])
        self.n = n
        self._keep = [np.ascontiguousarray(s["xyz"], np.int32), np.ascontiguousarray(s["sdf0"], np.float64),
                      np.ascontiguousarray(s["sdf_refined"], np.float64), np.ascontiguousarray(s["albedo"], np.float64),
                      np.ascontiguousarray(s["weight"], np.float32), np.ascontiguousarray(s["rgb"], np.uint8)]
        k = self._keep
        self._check(self.L.i3do_set_grid(self.h, C.c_int64(n), _p(k[0], C.c_int32), _p(k[1], C.c_double), _p(k[2], C.c_double),
                                         _p(k[3], C.c_double), _p(k[4], C.c_float), _p(k[5], C.c_uint8),
                                         C.c_float(float(s["voxel_size"]))))

    def estimate_lighting(self, params: I3DLightingParams) -> I3DLightingInfo:
        info = I3DLightingInfo()
        self._check(self.L.i3do_estimate_lighting(self.h, C.byref(params), C.byref(info)))
        return info

    def lighting(self):
        S = int(self.L.i3do_num_subvolumes(self.h))
        idx = np.empty((S, 3), np.int32)
        sh = np.empty((S, 9), np.float64)
        self.L.i3do_get_lighting(self.h, _p(idx, C.c_int32), _p(sh, C.c_double))
        return idx, sh

    def lighting_rows(self):
        """SHDataCost rows (subvolume, voxel, 9-column Jacobian row, target luminance, raw weight) and directed pairs."""
        m = int(self.L.i3do_num_lighting_rows(self.h, C.c_int(0)))
        npairs = int(self.L.i3do_num_lighting_rows(self.h, C.c_int(1)))
        sub = np.empty(m, np.int32)
        vox = np.empty(m, np.int32)
        j = np.empty((m, 9), np.float64)
        lum = np.empty(m, np.float64)
        w = np.empty(m, np.float64)
        pairs = np.empty((npairs, 2), np.int32)
        self.L.i3do_get_lighting_rows(self.h, _p(sub, C.c_int32), _p(vox, C.c_int32), _p(j, C.c_double), _p(lum, C.c_double), _p(w, C.c_double),
                                      _p(pairs, C.c_int32))
        return dict(sub=sub, voxel=vox, j=j, lum=lum, w=w, pairs=pairs)

    def voxel_sh(self):
        sh = np.empty((self.n, 9), np.float64)
        has = np.empty(self.n, np.uint8)
        self.L.i3do_get_voxel_sh(self.h, _p(sh, C.c_double), _p(has, C.c_uint8))
        return sh, has

    def step(self):
        U = 2 * self.n + 6 * self.F + 9
        st = np.zeros(U, np.float64)
        fm = np.zeros(U, np.uint8)
        cs = np.zeros(U, np.float64)
        self.L.i3do_get_step(self.h, _p(st, C.c_double), _p(fm, C.c_uint8), _p(cs, C.c_double))
        return st, fm, cs


def eval_eg(coord, voxel_size, pyr_scale, lum, sh, sdf, alb, pose, intr, dist, want_jac=True):
    """Standalone E_g residual (+ raw 29-column Jacobian) at explicit parameter values."""
    L = lib()
    lum = np.ascontiguousarray(lum, np.float32)
    h, w = lum.shape
    coord = np.ascontiguousarray(coord, np.int32)
    arrs = [np.ascontiguousarray(a, np.float64) for a in (sh, sdf, alb, pose, intr, dist)]
    res = C.c_double(0.0)
    jac = np.zeros(29, np.float64)
    L.i3do_eval_eg(_p(coord, C.c_int32), C.c_double(voxel_size), C.c_double(pyr_scale), C.c_int(w), C.c_int(h),
                   _p(lum, C.c_float), *[_p(a, C.c_double) for a in arrs], C.byref(res),
                   _p(jac, C.c_double) if want_jac else None)
    return res.value, jac


def default_lighting_params() -> I3DLightingParams:
    p = I3DLightingParams()
    lib().i3do_default_lighting_params(C.byref(p))
    return p
