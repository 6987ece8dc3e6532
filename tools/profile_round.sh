#!/bin/bash
# One-shot profile of the headline step on a B200 (run under gpurun): launch list, ncu --set full of the GEMM launches of
# one steady-state step and of the gather on 2^20 rows, per-tile GEMM timelines, the bench line.  Outputs -> gpurun_out/.
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py > $O/bench_1gpu_$R.json 2> $O/bench_1gpu_$R.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bf16_$R.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $O/ncu_launches_$R.log 2>&1
python tools/parse_launches.py $O/launches_bf16_$R.csv > $O/launches_bf16_$R.txt 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_grouped_tc2 --launch-skip 40 -c 10 -f \
    -o $O/prof_tc2_$R python bench.py --steps 2 --warmup 3 --no-extras > $O/ncu_tc2_$R.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gather_concat --launch-skip 3 -c 1 -f \
    -o $O/prof_gather_$R python tools/gather_bench.py > $O/ncu_gather_$R.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:"gate_level|heads_fast" --launch-skip 20 -c 7 -f \
    -o $O/prof_gate_heads_$R python bench.py --steps 2 --warmup 3 --no-extras > $O/ncu_gate_$R.log 2>&1
timeout 120 python tools/tc_timeline.py > $O/tc_timeline_$R.txt 2>&1
echo profile done
