import sys, time, numpy as np
sys.path.insert(0, '.')
from intrinsic3d_b200.scene import config_scene
from intrinsic3d_b200.ctypes_defs import default_params
from intrinsic3d_b200.engine import Engine
from oracle import Oracle
s = config_scene('small')
e = Engine(0); e.load_scene(s)
o = Oracle(); o.load_scene(s)
p = default_params(); p.thres_shell = s['thres_shell']
for forced in (3, 0):
    p.forced_cg_iterations = forced
    e.upload_voxel_params(s['sdf_refined'], s['albedo']); e.set_camera(s['poses'], s['intr'], s['dist'])
    o.load_scene(s)
    t=time.time(); ie = e.gn_iteration(p); te=time.time()-t
    t=time.time(); io = o.gn_iteration(p); to=time.time()-t
    print('forced', forced, 'gpu s', te, 'cpu s', to)
    for k in ['num_active','num_free_sdf','num_parameters','type_residuals','type_sum_weights','type_weights','type_costs','cost_initial','cost_final','trust_region_radius','lm_iterations','step_accepted','termination','cg_iterations_total','step_norm']:
        a=getattr(ie,k); b=getattr(io,k)
        if hasattr(a,'__len__'): a=list(a); b=list(b)
        print(' ',k, a, b)
    n=ie.lm_iterations
    print('  cg', list(ie.cg_iterations)[:n], list(io.cg_iterations)[:n])
    print('  mcc', list(ie.model_cost_change)[:n], list(io.model_cost_change)[:n])
    print('  cand', list(ie.candidate_cost)[:n], list(io.candidate_cost)[:n])
    se,_,cse = e.debug_step(); so,fmo,cso = o.step()
    print('  step maxabs diff', np.abs(se-so).max(), 'ref', np.abs(so).max())
    for ph in ['total','select','build','solve','pcg','candidate']:
        print('  phase', ph, e.phase_ms(ph), e.phase_count(ph))
