#!/bin/bash
# Final multi-GPU evidence run of a round (run under `gpurun --gpus N`):  bash profiles/mg8_final.sh <tag> <N>
# sharded-vs-unsharded parity with both transports, bench.py at N GPUs with the row-balanced and the time-balanced shards, C5 end to end.
TAG="$1"; N="$2"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
I3D_MG_SCENE=small I3D_MG_BALANCE=1 timeout 150 $TR tests/mg_check.py > gpurun_out/${TAG}_mgcheck_p2p_n$N.log 2>&1; tail -2 gpurun_out/${TAG}_mgcheck_p2p_n$N.log
I3D_MG_SCENE=small I3D_XCHG=nccl timeout 150 $TR tests/mg_check.py > gpurun_out/${TAG}_mgcheck_nccl_n$N.log 2>&1; tail -1 gpurun_out/${TAG}_mgcheck_nccl_n$N.log
timeout 200 $TR bench.py --gpus $N --steps 25 --warmup 5 --no-e2e > gpurun_out/${TAG}_bench_p2p_n$N.json 2> gpurun_out/${TAG}_bench_p2p_n$N.err
I3D_SHARD=timed timeout 200 $TR bench.py --gpus $N --steps 25 --warmup 5 --no-e2e --no-selfcheck > gpurun_out/${TAG}_bench_p2p_timed_n$N.json 2> gpurun_out/${TAG}_bench_p2p_timed_n$N.err
I3D_C5_MAX_STEPS=1 timeout 240 $TR bench.py --gpus $N --workload c5 --steps 1 > gpurun_out/${TAG}_bench_c5_n$N.json 2> gpurun_out/${TAG}_bench_c5_n$N.err
python - "$TAG" "$N" <<'PY'
import json, sys, statistics as st
tag, n = sys.argv[1], sys.argv[2]
for t in ("p2p", "p2p_timed"):
    try:
        d = json.loads(open(f"gpurun_out/{tag}_bench_{t}_n{n}.json").read().strip().splitlines()[-1])
        ph = d["per_step"]["phase_ms"]
        print(t, "iter/s %.1f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(st.median(v), 3) for k, v in ph.items()}, "gap", d["per_step"].get("host_gap_ms_mean"),
              "selfcheck", (d.get("mg_selfcheck") or {}).get("ok"), d["per_step"]["kernel_ms_detail_step"])
        print("   per-rank:", d["per_step"].get("per_rank_mean_ms"))
        print("   shards:", d.get("shard_ranges"))
    except Exception as ex:
        print(t, "not run or failed:", ex)
try:
    d = json.loads(open(f"gpurun_out/{tag}_bench_c5_n{n}.json").read().strip().splitlines()[-1])
    print("c5", "iter/s %.1f ms/refine %.1f" % (d["value"], d["ms_per_step"]), [(l["voxels"], round(l["wall_s"], 3)) for l in d["levels"]])
except Exception as ex:
    print("c5 not run or failed:", ex)
    try:
        print(open(f"gpurun_out/{tag}_bench_c5_n{n}.err").read()[-1500:])
    except Exception:
        pass
PY
