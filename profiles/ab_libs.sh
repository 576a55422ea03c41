#!/bin/bash
# A/B of alternative builds of the engine library on one GPU box:  gpurun -- 'bash profiles/ab_libs.sh <tag> lib1.so lib2.so ...'
# (I3D_LIB selects the library the Python binding loads; each run = short resident-only bench, kernel times from CUDA events)
TAG="$1"; shift
mkdir -p gpurun_out
for L in "$@"; do
  B=$(basename "$L" .so)
  I3D_LIB="$PWD/$L" timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check --no-lighting > "gpurun_out/${TAG}_$B.json" 2> "gpurun_out/${TAG}_$B.err"
  python - "$B" "gpurun_out/${TAG}_$B.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    ks=d["per_step"]["kernel_ms_detail_step"]
    print(sys.argv[1], "iter/s %.1f" % d["value"], "build %.3f cost %.3f apply %.3f/launch select %.3f accum %.3f syncs %s gap %s" % (ks["k_eg_build"][0], ks["k_eg_cost"][0]/max(1,ks["k_eg_cost"][1]), ks["k_eg_apply"][0]/max(1,ks["k_eg_apply"][1]), ks["k_select_obs"][0], ks["k_eg_accum"][0], d.get("host_syncs_per_step"), d["per_step"].get("host_gap_ms_mean")))
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
