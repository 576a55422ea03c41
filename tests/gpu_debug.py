"""Debug helper (run under compute-sanitizer on a GPU box): one full GN iteration on a z-slab of a scene (missing stencil neighbours at the cut,
odd voxel count).  usage: python tests/gpu_debug.py [scene] [fraction]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from bench import make_params, slab_scene
    from intrinsic3d_b200.engine import Engine
    from intrinsic3d_b200.scene import config_scene
    name = sys.argv[1] if len(sys.argv) > 1 else "small"
    frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
    import torch
    s, m = slab_scene(config_scene(name, device="cuda" if torch.cuda.is_available() else "cpu"), frac)
    if m % 2 == 0:
        for k in ("xyz", "sdf0", "sdf_refined", "albedo", "weight", "rgb", "sh"):
            s[k] = s[k][:-1]
    e = Engine(0)
    e.load_scene(s)
    p = make_params(s)
    for it in range(2):
        info = e.gn_iteration(p)
        print("iteration", it, "cg", list(info.cg_iterations)[:info.lm_iterations], "accepted", info.step_accepted, "cost", info.cost_initial, info.cost_final, flush=True)


if __name__ == "__main__":
    main()
