#!/bin/bash
# On the GPU box: tests, bench line, ncu launch list and full captures of this round.  Nothing printed under ncu is a bench value.
cd "$(dirname "$0")/.." || exit 1
R=${1:-r02}
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputest.log 2>&1; echo "rc=$?" >> gpurun_out/${R}_gputest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench.log 2>&1 && tail -1 gpurun_out/${R}_bench.log > gpurun_out/${R}_bench_line.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${R}_launch_list.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${R}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_pass -s 2 -c 1 -o gpurun_out/${R}_fused_pass python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${R}_ncu_fused.log 2>&1
FSUAE_NO_MEGA=1 timeout 600 ncu --set full --clock-control none -k regex:conv3x3_tc -s 12 -c 6 -o gpurun_out/${R}_layers python bench.py --steps 2 --warmup 3 --quick > gpurun_out/${R}_ncu_layers.log 2>&1
tail -3 gpurun_out/${R}_gputest.log; tail -c 600 gpurun_out/${R}_bench_line.json
