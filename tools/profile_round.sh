#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the bench command + one `ncu --set full` capture of K1 and of K2.
# Usage: tools/profile_round.sh <tag>          (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-prof}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 20 --warmup 3 --no-cpu-baseline"
# (1) every launch of the bench command with its device time (cold cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
    $BENCH > $OUT/${TAG}_launches_bench.log 2>&1
# (2) K1 (headline kernel), full set, two launches after the warm-up
ncu --set full --clock-control none --import-source on -k regex:k1_top2 -s 3 -c 2 -f -o $OUT/${TAG}_k1 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $OUT/${TAG}_k1.log 2>&1
# (2b) K1T (tensor-core engine, the headline): expansion, search and finish kernels of two steps after its warm-up
#      -> python tools/ncu_summary.py roofline_k1t $OUT/${TAG}_k1t.ncu-rep profiles/k1t_roofline.json   (read here)
ncu --set full --clock-control none --import-source on -k regex:k1t -s 9 -c 6 -f -o $OUT/${TAG}_k1t \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary --no-survey > $OUT/${TAG}_k1t.log 2>&1
# (3) K2 (MSAC scoring, configs[2]), full set
ncu --set full --clock-control none --import-source on -k regex:k2_score_kernel -s 1 -c 2 -f -o $OUT/${TAG}_k2 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_k2.log 2>&1
ls -la $OUT
