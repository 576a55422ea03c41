import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, '.')
from intrinsic3d_b200.scene import config_scene
from intrinsic3d_b200.ctypes_defs import default_params
from intrinsic3d_b200.engine import Engine
s = config_scene('small'); its=2
lam = np.array([0.2, 80.0, 10.0, 120.0, 10.0, 0.1])
e = Engine(0); e.load_scene(s)
for it in range(its):
    p = default_params(); p.thres_shell = s["thres_shell"]
    p.lambda_[0]=lam[0]; p.lambda_[1]=lam[1]+(lam[2]-lam[1])/(its-1)*it; p.lambda_[2]=lam[3]+(lam[4]-lam[3])/(its-1)*it; p.lambda_[3]=lam[5]
    i=e.gn_iteration(p); print('py it',it,list(i.type_residuals), i.cost_initial, i.cost_final, i.cg_iterations_total, list(p.lambda_))
ref=e.download_state()
def P(a,t): return a.ctypes.data_as(C.POINTER(t))
H = C.CDLL(os.path.abspath("intrinsic3d_b200/libi3d_host.so"))
n=s["xyz"].shape[0]; F,Hh,W=s["lum"].shape
xyz=np.ascontiguousarray(s["xyz"],np.int32); sdf0=np.ascontiguousarray(s["sdf0"],np.float64); sdf=s["sdf_refined"].astype(np.float64).copy(); alb=s["albedo"].astype(np.float64).copy()
wgt=np.ascontiguousarray(s["weight"],np.float32); rgb=np.ascontiguousarray(s["rgb"],np.uint8); lum=np.ascontiguousarray(s["lum"],np.float32); dep=np.ascontiguousarray(s["depth"],np.float32)
poses=s["poses"].astype(np.float64).copy(); intr=s["intr"].astype(np.float64).copy(); dist=s["dist"].astype(np.float64).copy(); sh=np.ascontiguousarray(s["sh"],np.float64); counts=np.zeros(4,np.int64)
os.environ["I3D_HOST_DEBUG"]="1"
rc=H.i3dh_run_optimizer(C.c_int64(n),P(xyz,C.c_int32),P(sdf0,C.c_double),P(sdf,C.c_double),P(alb,C.c_double),P(wgt,C.c_float),P(rgb,C.c_uint8),C.c_float(float(s["voxel_size"])),C.c_int32(F),C.c_int32(W),C.c_int32(Hh),P(lum,C.c_float),P(dep,C.c_float),P(poses,C.c_double),P(intr,C.c_double),P(dist,C.c_double),P(sh,C.c_double),C.c_double(s["thres_shell"]),C.c_float(0.02),C.c_int32(5),C.c_int32(its),C.c_int32(50),P(lam,C.c_double),C.c_int32(0),C.c_int32(0),C.c_int32(0),P(counts,C.c_int64))
print('rc',rc,counts, np.abs(sdf-ref["sdf_refined"]).max(), np.abs(poses-ref["poses"]).max(), np.abs(intr-ref["intr"]).max())
