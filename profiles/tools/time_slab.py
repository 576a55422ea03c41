"""Kernel times of one GN iteration on a z-slab (fraction of the C3 grid, all frames) on ONE GPU: what a rank of an N-GPU run sees
(single-wave launches).  python profiles/tools/time_slab.py <fraction>"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from bench import make_params, lambda_schedule, slab_scene
from intrinsic3d_b200.engine import Engine
from intrinsic3d_b200.scene import config_scene
frac = float(sys.argv[1]) if len(sys.argv) > 1 else 0.125
scene = config_scene("c3", device="cuda:0")
s, m = slab_scene(scene, frac)
eng = Engine(0); eng.load_scene(s); p = make_params(scene)
eng.set_kernel_timers(1)
out = []
for it in range(6):
    lambda_schedule(p, it); eng.gn_iteration(p)
    out.append({k: round(eng.phase_ms(k), 4) for k in ("k_select_obs", "k_eg_build", "k_eg_cost", "k_eg_accum", "k_eg_apply", "select", "total")})
print(os.environ.get("I3D_LIB", "default"), "voxels", m, out[2:])
