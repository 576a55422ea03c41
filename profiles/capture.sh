#!/bin/bash
# Captures the ncu evidence for one round on a GPU box (run under gpurun, ONE GPU):
#   gpurun --timeout 900 -- 'bash profiles/capture.sh r02'
# 1) launch list of one bench step (shares), 2) --set full reports of the kernels named below, written to gpurun_out/ (scratch);
# then, back in the build container:  python profiles/summarize.py <tag> gpurun_out/pf_*.ncu-rep   (extracts the numbers into profiles/)
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file "$OUT/${TAG}_launches.csv" $BENCH > "$OUT/${TAG}_launches.log" 2>&1
for K in k_eg_apply k_eg_build k_eg_cost k_select_obs k_eg_accum k_svsh_accumulate k_svsh_solve k_recolor k_upsample k_shell_crossing; do
    # one launch of each kernel, skipping the warm-up launches of the per-iteration kernels
    SKIP=0
    case "$K" in k_eg_apply) SKIP=40 ;; k_eg_build|k_eg_cost|k_select_obs|k_eg_accum) SKIP=3 ;; esac
    timeout 300 ncu --set full --clock-control none --import-source on -k "regex:^$K" -s $SKIP -c 1 -f -o "$OUT/pf_$K" $BENCH > "$OUT/pf_$K.log" 2>&1 || echo "capture of $K failed (see $OUT/pf_$K.log)"
done
ls -la "$OUT"/pf_*.ncu-rep 2>/dev/null
