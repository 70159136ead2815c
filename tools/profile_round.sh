set -x
timeout 500 python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r01_bench_line.json
timeout 300 python tools/bench_models.py > gpurun_out/r01_model_families.log 2>&1
for a in "conv3 heavyweight" "conv5 heavyweight" "pix_shuffle heavyweight" "conv3 lightweight" "conv5 lightweight"; do echo "== $a"; timeout 120 python tools/kernel_times.py $a bf16 16; done > gpurun_out/r01_family_kernel_times.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 16 --csv --log-file gpurun_out/r01_launch_list.csv python bench.py --steps 2 --warmup 3 --quick > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc_wide -s 2 -c 2 -o gpurun_out/wide_conv3heavy python tools/kernel_times.py conv3 heavyweight bf16 8 > gpurun_out/ncu_wide.log 2>&1
tail -2 gpurun_out/ncu_wide.log
