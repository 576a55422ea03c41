#!/bin/bash
# One development step on a GPU box (run under gpurun, ONE GPU):
#   gpurun --timeout 900 -- 'bash profiles/gpu_step.sh <tag> [pytest-args]'
# core parity tests -> short bench (resident legs only) -> one `ncu --set full` capture of the E_g row kernels.
# Everything lands in gpurun_out/ (scratch); `python profiles/summarize.py <tag> gpurun_out/pf_*.ncu-rep` extracts the numbers.
set -u
TAG="${1:-rXX}"
TESTS="${2:-tests/test_gpu_parity.py}"
OUT=gpurun_out
mkdir -p "$OUT"
timeout 600 python -m pytest $TESTS -x -q > "$OUT/${TAG}_tests.log" 2>&1
tail -5 "$OUT/${TAG}_tests.log"
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-lighting > "$OUT/${TAG}_bench.json" 2> "$OUT/${TAG}_bench.err"
tail -c 1500 "$OUT/${TAG}_bench.json"; tail -5 "$OUT/${TAG}_bench.err"
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-lighting"
for K in ${NCU_KERNELS:-k_eg_rows}; do
    SKIP=6; CNT=2
    case "$K" in k_eg_apply) SKIP=40; CNT=1 ;; k_select_obs|k_eg_accum) SKIP=3; CNT=1 ;; esac
    timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$K" -s $SKIP -c $CNT -f -o "$OUT/pf_${TAG}_$K" $BENCH > "$OUT/pf_${TAG}_$K.log" 2>&1 || echo "capture of $K failed"
done
ls -la "$OUT" | tail -20
