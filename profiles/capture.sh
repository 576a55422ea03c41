#!/bin/bash
# Captures the ncu / sanitizer evidence for one round on a GPU box (run under gpurun, ONE GPU):
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r02'
# 1) launch list of one bench step (shares), 2) --set full reports of the kernels named below, 3) compute-sanitizer racecheck + memcheck
# on the small scene; everything goes to gpurun_out/ (scratch).  Back in the build container:
#   python profiles/summarize.py <tag> gpurun_out/pf_<tag>_*.ncu-rep      (extracts the numbers into profiles/)
set -u
TAG="${1:-rXX}"
OUT=gpurun_out
mkdir -p "$OUT"
BENCH="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity-check ${BENCH_EXTRA:-}"
# launch list of the ENGINE's kernels only (the synthetic scene generator launches ~700 torch kernels first): one warm-up + one step
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base function -k "regex:^k_" -c 600 --csv --log-file "$OUT/${TAG}_launches.csv" $BENCH > "$OUT/${TAG}_launches.log" 2>&1
for K in ${NCU_KERNELS:-k_eg_apply k_eg_rows k_select_obs k_eg_accum k_op_partial k_svsh_accumulate k_svsh_solve k_recolor k_upsample k_shell_crossing}; do
    # one or two launches of each kernel, skipping the warm-up launches of the per-iteration kernels
    SKIP=0; CNT=1
    case "$K" in k_eg_apply|k_op_partial) SKIP=40 ;; k_eg_rows) SKIP=6; CNT=2 ;; k_select_obs|k_eg_accum) SKIP=3 ;; esac
    timeout 400 ncu --set full --clock-control none --import-source on -k "regex:^$K" -s $SKIP -c $CNT -f -o "$OUT/pf_${TAG}_$K" $BENCH > "$OUT/pf_${TAG}_$K.log" 2>&1 || echo "capture of $K failed (see $OUT/pf_${TAG}_$K.log)"
done
if [ "${SANITIZE:-1}" = "1" ]; then
    SAN="python -c \"import __graft_entry__ as g; g.smoke()\""
    timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/${TAG}_racecheck.log" 2>&1; tail -3 "$OUT/${TAG}_racecheck.log"
    timeout 900 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/${TAG}_memcheck.log" 2>&1; tail -3 "$OUT/${TAG}_memcheck.log"
fi
ls -la "$OUT"/pf_${TAG}_*.ncu-rep 2>/dev/null
