#!/bin/bash
# Multi-GPU development step (run under `gpurun --gpus N`):  bash profiles/mg_step.sh <tag> <N> [scene]
# 1) sharded-vs-unsharded parity (tests/mg_check.py) with the peer-memory exchange + balanced shards, and again with ncclAllReduce,
# 2) bench.py at N GPUs (resident legs + mg_selfcheck).  Logs -> gpurun_out/.
TAG="$1"; N="$2"; SCENE="${3:-small}"
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555"
I3D_MG_SCENE=$SCENE I3D_MG_BALANCE=1 timeout 600 $TR tests/mg_check.py > gpurun_out/${TAG}_mgcheck_p2p_n$N.log 2>&1; tail -6 gpurun_out/${TAG}_mgcheck_p2p_n$N.log
I3D_MG_SCENE=$SCENE I3D_XCHG=nccl timeout 600 $TR tests/mg_check.py > gpurun_out/${TAG}_mgcheck_nccl_n$N.log 2>&1; tail -3 gpurun_out/${TAG}_mgcheck_nccl_n$N.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e > gpurun_out/${TAG}_bench_p2p_n$N.json 2> gpurun_out/${TAG}_bench_p2p_n$N.err
if [ "${MG_NCCL_BENCH:-0}" = "1" ]; then I3D_XCHG=nccl timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-e2e --no-selfcheck > gpurun_out/${TAG}_bench_nccl_n$N.json 2> gpurun_out/${TAG}_bench_nccl_n$N.err; fi
python - "$TAG" "$N" <<'PY'
import json,sys
for t in ("p2p","nccl"):
    try:
        d=json.loads(open(f"gpurun_out/{sys.argv[1]}_bench_{t}_n{sys.argv[2]}.json").read().strip().splitlines()[-1])
        ks=d["per_step"]["kernel_ms_detail_step"]; ph=d["per_step"]["phase_ms"]
        import statistics as st
        print(t, "iter/s %.1f ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(st.mean(v),3) for k,v in ph.items()}, "gap", d["per_step"].get("host_gap_ms_mean"), "selfcheck", (d.get("mg_selfcheck") or {}).get("ok"), ks)
    except Exception as ex:
        print(t, "not run or failed:", ex)
PY
