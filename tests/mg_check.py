"""Multi-GPU parity driver (run under torchrun, one rank per GPU): the voxel-sharded engine must reproduce the
single-GPU engine on the same scene.  Invoked by tests/test_gpu_multi.py."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from intrinsic3d_b200.ctypes_defs import default_params
    from intrinsic3d_b200.engine import Engine, shard_range
    from intrinsic3d_b200.scene import config_scene

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s = config_scene(os.environ.get("I3D_MG_SCENE", "small"))
    n = s["xyz"].shape[0]
    e = Engine(local)
    e.comm_init(rank, world, dist)
    e.load_scene(s)
    if os.environ.get("I3D_MG_BALANCE", "0") == "1":
        from intrinsic3d_b200.engine import balanced_shard_ranges
        p0 = default_params()
        p0.thres_shell = s["thres_shell"]
        ranges = balanced_shard_ranges(e, dist, p0, n)
        e.set_shard(*ranges[rank])
        if rank == 0:
            print("balanced shard ranges:", ranges, flush=True)
    else:
        e.set_shard(*shard_range(n, rank, world, align=64))
    if rank == 0:
        print(f"exchange transport: {'peer memory (NVLink pulls)' if getattr(e, 'p2p', False) else 'ncclAllReduce'}", flush=True)
    ref = Engine(local)          # same GPU, unsharded
    ref.load_scene(s)
    ok = True
    for it in range(3):
        p = default_params()
        p.thres_shell = s["thres_shell"]
        p.lambda_[1] = 80.0 - 70.0 / 9.0 * it
        p.lambda_[2] = 120.0 - 110.0 / 9.0 * it
        a = e.gn_iteration(p)
        b = ref.gn_iteration(p)
        sa, sb = e.download_state(), ref.download_state()
        step = ref.debug_step()[0]
        refmag = np.abs(step[:n]).max()
        checks = dict(
            rows=list(a.type_residuals) == list(b.type_residuals),
            active=a.num_active == b.num_active,
            sums=np.allclose(list(a.type_sum_weights), list(b.type_sum_weights), rtol=1e-9),
            cost0=abs(a.cost_initial - b.cost_initial) <= 1e-9 * abs(b.cost_initial),
            cg=list(a.cg_iterations)[:a.lm_iterations] == list(b.cg_iterations)[:b.lm_iterations],
            accepted=a.step_accepted == b.step_accepted,
            cost1=abs(a.cost_final - b.cost_final) <= 1e-5 * abs(b.cost_final),
            sdf=np.abs(sa["sdf_refined"] - sb["sdf_refined"]).max() <= 1e-3 * refmag,
            albedo=np.abs(sa["albedo"] - sb["albedo"]).max() <= 1e-3 * max(np.abs(step[n:2 * n]).max(), 1e-30),
            poses=np.abs(sa["poses"] - sb["poses"]).max() <= 1e-3 * max(np.abs(step[2 * n:2 * n + 6 * s["poses"].shape[0]]).max(), 1e-30),
        )
        if rank == 0:
            print(f"iter {it}: cg {list(a.cg_iterations)[:a.lm_iterations]} vs {list(b.cg_iterations)[:b.lm_iterations]} "
                  f"cost {a.cost_final:.9f} vs {b.cost_final:.9f} checks {checks}", flush=True)
        ok = ok and all(checks.values())
        # keep every engine on identical inputs: the unsharded reference runs once per rank and (atomics) is not bitwise
        # reproducible between GPUs, so rank 0's state is broadcast
        for k in ("sdf_refined", "albedo", "poses", "intr", "dist"):
            t = torch.from_numpy(sb[k]).cuda()
            dist.broadcast(t, 0)
            sb[k] = t.cpu().numpy()
        for eng in (e, ref):
            eng.upload_voxel_params(sb["sdf_refined"], sb["albedo"])
            eng.set_camera(sb["poses"], sb["intr"], sb["dist"])
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MG_CHECK_OK" if t.item() == 1 else "MG_CHECK_FAIL", flush=True)
    sys.exit(0 if t.item() == 1 else 1)


if __name__ == "__main__":
    main()
