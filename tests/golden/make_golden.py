"""Generates tests/golden/tiny_gn.npz: a small seeded scene (inputs included, so the fixture does not depend on the
generator's device or torch version) and the CPU oracle's outputs for one Gauss-Newton iteration at a fixed PCG
iteration count.  The reference ships no golden vectors; these pin the oracle against regressions and
give the GPU parity tests a committed target.   Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

FORCED_CG = 6


def main():
    from intrinsic3d_b200.ctypes_defs import default_params
    from intrinsic3d_b200.scene import make_scene
    from oracle import Oracle
    s = make_scene(radius_vox=7.0, frames=5, width=120, height=90, seed=5)
    o = Oracle(threads=4)
    o.load_scene(s)
    p = default_params()
    p.thres_shell = s["thres_shell"]
    p.forced_cg_iterations = FORCED_CG
    info = o.gn_iteration(p)
    K = 5
    fr, w, act = o.observations(K)
    eg = o.rows(0)
    step, free, scale = o.step()
    st = o.state()
    out = dict(
        # inputs
        xyz=s["xyz"], sdf0=s["sdf0"], sdf_refined=s["sdf_refined"], albedo=s["albedo"], weight=s["weight"], rgb=s["rgb"],
        voxel_size=np.float32(s["voxel_size"]), lum=s["lum"].astype(np.float16).astype(np.float32), depth=s["depth"],
        poses=s["poses"], intr=s["intr"], dist=s["dist"], sh=s["sh"], thres_shell=np.float64(s["thres_shell"]),
    )
    # the luminance is stored as float16-representable values to keep the fixture small: regenerate the outputs on exactly these inputs
    s2 = dict(s)
    s2["lum"] = out["lum"]
    o = Oracle(threads=4)
    o.load_scene(s2)
    info = o.gn_iteration(p)
    fr, w, act = o.observations(K)
    eg = o.rows(0)
    step, free, scale = o.step()
    st = o.state()
    out.update(
        forced_cg=np.int32(FORCED_CG), obs_frames=fr, obs_weights=w, active=act,
        eg_voxel=eg["voxel"], eg_frame=eg["aux"], eg_residual=eg["residual"], eg_raw_weight=eg["raw_weight"],
        eg_jacobian=o.eg_jacobian().astype(np.float32),
        type_residuals=np.array(list(info.type_residuals)), type_sum_weights=np.array(list(info.type_sum_weights)),
        type_costs=np.array(list(info.type_costs)), cost_initial=np.float64(info.cost_initial), cost_final=np.float64(info.cost_final),
        model_cost_change=np.float64(info.model_cost_change[0]), lm_iterations=np.int32(info.lm_iterations), accepted=np.int32(info.step_accepted),
        trust_region_radius=np.float64(info.trust_region_radius), step=step.astype(np.float32), free_mask=free,
        out_sdf=st["sdf_refined"], out_albedo=st["albedo"], out_poses=st["poses"], out_intr=st["intr"], out_dist=st["dist"],
    )
    path = os.path.join(ROOT, "tests", "golden", "tiny_gn.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; rows", list(info.type_residuals), "cost", info.cost_initial, "->", info.cost_final)
    make_lighting(out)
    make_recolor(out)
    make_gridops(out)


RECOLOR_K = 2


def make_recolor(inputs):
    """tests/golden/tiny_recolor.npz: colour frames (B,G,R) for the scene of tiny_gn.npz and the oracle's
    Intrinsic3D::recomputeColors result at K = 2 (so that the top-K filter runs) and K = 0."""
    import oracle
    from intrinsic3d_b200.scene import make_color_frames
    s = {k: inputs[k] for k in ("xyz", "sdf0", "sdf_refined", "albedo", "weight", "rgb", "lum", "depth", "poses", "intr", "dist", "sh")}
    s["voxel_size"] = inputs["voxel_size"]
    col = make_color_frames(s, seed=11)
    out = dict(color=col, K=np.int32(RECOLOR_K), occlusion=np.float32(0.02))
    for tag, K in (("k", RECOLOR_K), ("all", 0)):
        o = oracle.Oracle(threads=2)
        o.load_scene(s)
        o.set_color_frames(col)
        cnt = o.recompute_colors(0.02, K)
        out["rgb_" + tag] = o.colors()
        out["counts_" + tag] = np.array(cnt, np.int64)
    path = os.path.join(ROOT, "tests", "golden", "tiny_recolor.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", out["counts_k"], out["counts_all"], int((out["rgb_k"] != out["rgb_all"]).any(1).sum()), "voxels differ between K=2 and all")


def make_gridops(inputs):
    """tests/golden/tiny_gridops.npz: the oracle's clearVoxelsOutsideThinShell (shell = 1 voxel) on the grid of tiny_gn.npz with 60 voxels
    invalidated, followed by upsample and a second pruning at the fine level (a grid-level transition of Intrinsic3D::refine)."""
    import oracle
    s = {k: inputs[k] for k in ("xyz", "sdf0", "sdf_refined", "albedo", "weight", "rgb", "lum", "depth", "poses", "intr", "dist", "sh")}
    s["voxel_size"] = inputs["voxel_size"]
    w = s["weight"].copy()
    w[::53] = 0.0
    s["weight"] = w
    vs = float(np.float32(s["voxel_size"]))
    o = oracle.Oracle(threads=2)
    o.load_scene(s)
    out = dict(weight_in=w, shell=np.float64(vs))
    m1 = o.clear_voxels_outside_thin_shell(vs)
    g1 = o.grid()
    m2 = o.upsample_grid()
    g2 = o.grid()
    m3 = o.clear_voxels_outside_thin_shell(0.5 * vs)
    g3 = o.grid()
    out.update(counts=np.array([m1, m2, m3], np.int64), prune_xyz=g1["xyz"], prune_sdf=g1["sdf_refined"],
               up_xyz_head=g2["xyz"][:4096], up_sdf=g2["sdf_refined"].astype(np.float32), up_sdf0=g2["sdf0"].astype(np.float32),
               up_albedo=g2["albedo"].astype(np.float32), up_weight=g2["weight"], up_rgb=g2["rgb"], up_voxel_size=g2["voxel_size"],
               final_xyz=g3["xyz"], final_rgb=g3["rgb"])
    path = os.path.join(ROOT, "tests", "golden", "tiny_gridops.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", out["counts"])


LIGHT_SUBVOLUME_SIZE = 0.012
LIGHT_LAMBDA_REG = 10.0


def make_lighting(inputs):
    """tests/golden/tiny_lighting.npz: the oracle's LightingSVSH::estimate + computeVoxelShCoeffs on the grid of tiny_gn.npz."""
    import oracle
    s = {k: inputs[k] for k in ("xyz", "sdf0", "sdf_refined", "albedo", "weight", "rgb")}
    s["voxel_size"] = inputs["voxel_size"]
    o = oracle.Oracle(threads=2)
    o.set_grid(s)
    P = oracle.default_lighting_params()
    P.thres_shell = float(inputs["thres_shell"]); P.subvolume_size = LIGHT_SUBVOLUME_SIZE; P.lambda_reg = LIGHT_LAMBDA_REG
    info = o.estimate_lighting(P)
    idx, sh = o.lighting()
    vsh, has = o.voxel_sh()
    out = dict(subvolume_size=np.float32(LIGHT_SUBVOLUME_SIZE), lambda_reg=np.float64(LIGHT_LAMBDA_REG), sub_index=idx, sub_sh=sh,
               voxel_sh=vsh, has_sh=has, num_data_rows=np.int64(info.num_data_rows), num_reg_pairs=np.int64(info.num_reg_pairs),
               sum_data_weights=np.float64(info.sum_data_weights), cost_initial=np.float64(info.cost_initial),
               cost_final=np.float64(info.cost_final), lm_iterations=np.int32(info.lm_iterations),
               num_successful_steps=np.int32(info.num_successful_steps), cg_iterations_total=np.int32(info.cg_iterations_total),
               termination=np.int32(info.termination))
    path = os.path.join(ROOT, "tests", "golden", "tiny_lighting.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", info.as_dict())


if __name__ == "__main__":
    main()
