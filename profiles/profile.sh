#!/bin/bash
# Evidence run of a round (under gpurun, ONE GPU): launch list of one timed C2 step, `ncu --set full` captures of
# the hot kernels of charge 2, and a compute-sanitizer pass over the tiny workload. Everything lands in gpurun_out/;
# profiles/export_ncu.py turns the .ncu-rep files into the CSVs / traffic.json that are committed.
#   gpurun --timeout 1500 -- 'bash profiles/profile.sh r2'
TAG=${1:-r2c}
OUT=gpurun_out
mkdir -p $OUT
export SOLO_NCU=1
# 1. every launch of one step with its device time (cold cache, serialised: shares, not absolutes)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_launches.json 2> $OUT/${TAG}_launches.err
# 2. full counter sets of the hot kernels (first launches of the step = charge 2)
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:'scan_tc_kernel|k5_fast_kernel|select_probes|final_topk_kernel|threshold_kernel|coarse_tau' -c 14 \
    -o $OUT/${TAG}_kernels -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_kernels.json 2> $OUT/${TAG}_kernels.err
unset SOLO_NCU
# 3. memcheck + racecheck of the whole hot path on the tiny workload (mbarrier-heavy K3 included)
compute-sanitizer --tool memcheck --error-exitcode 1 python bench.py --workload tiny --steps 1 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?" >> $OUT/${TAG}_memcheck.log
compute-sanitizer --tool racecheck --error-exitcode 1 python bench.py --workload tiny --steps 1 --warmup 1 --no-cpu-baseline \
    > $OUT/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?" >> $OUT/${TAG}_racecheck.log
tail -3 $OUT/${TAG}_memcheck.log $OUT/${TAG}_racecheck.log
ls -la $OUT/${TAG}_*
