"""Small multi-iteration run used under compute-sanitizer (memcheck) on the GPU box."""
import sys
sys.path.insert(0, '.')
from intrinsic3d_b200.scene import config_scene
from intrinsic3d_b200.ctypes_defs import default_params
from intrinsic3d_b200.engine import Engine
s = config_scene('tiny')
e = Engine(0); e.load_scene(s)
p = default_params(); p.thres_shell = s['thres_shell']
for it in range(3):
    p.lambda_[1] = 80.0 - 7.0 * it
    i = e.gn_iteration(p)
    print('it', it, list(i.type_residuals), i.cost_initial, i.cost_final, i.cg_iterations_total, i.step_accepted)
p.forced_cg_iterations = 12      # exercises the exact-residual refresh path
i = e.gn_iteration(p); print('forced', i.cost_final, i.cg_iterations_total)
st = e.download_state(); print('ok', float(abs(st['sdf_refined']).max()))
