"""Times every stage of bench.py's e2e.refine_level_call separately (wall clock, device synchronised between stages)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from bench import make_params, lambda_schedule, ITERATIONS, color_frames_once
from intrinsic3d_b200.engine import Engine, default_lighting_params
from intrinsic3d_b200.scene import config_scene
scene = config_scene("c3", device="cuda:0")
eng = Engine(0)
eng.load_scene(scene)
p = make_params(scene)
keys = ("xyz", "sdf0", "sdf_refined", "albedo", "weight", "rgb", "lum", "depth", "poses", "intr", "dist", "sh")
host = {k: torch.from_numpy(np.ascontiguousarray(scene[k]).copy()).pin_memory().numpy() for k in keys}
col = torch.from_numpy(color_frames_once(scene)).pin_memory().numpy()
LP = default_lighting_params(); LP.thres_shell = scene["thres_shell"]
def stage(name, fn, log):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize(); log.append((name, 1e3 * (time.perf_counter() - t))); return r
for rep in range(3):
    log = []
    stage("upload_grid", lambda: eng.upload_grid(host["xyz"], host["sdf0"], host["sdf_refined"], host["albedo"], host["weight"], host["rgb"], scene["voxel_size"]), log)
    stage("upload_frames", lambda: eng.upload_frames(host["lum"], host["depth"], 1.0), log)
    stage("upload_color", lambda: eng.upload_color_frames(col), log)
    stage("set_camera", lambda: eng.set_camera(host["poses"], host["intr"], host["dist"]), log)
    stage("prune", lambda: eng.clear_voxels_outside_thin_shell(float(scene["thres_shell"])), log)
    stage("lighting", lambda: eng.estimate_lighting(LP), log)
    for it in range(ITERATIONS):
        lambda_schedule(p, it)
        stage(f"gn{it}", lambda: eng.gn_iteration(p), log)
    stage("recolor", lambda: eng.recompute_colors(0.02, 5), log)
    stage("download_grid", lambda: eng.download_grid(), log)
    stage("download_state", lambda: eng.download_state(), log)
    print(rep, "total %.1f ms" % sum(v for _, v in log), " ".join(f"{k}={v:.1f}" for k, v in log))
