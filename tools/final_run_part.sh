set -x
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err; tail -c 200 gpurun_out/r2_bench.log
python bench.py --domains 3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_d3.log 2> gpurun_out/r2_bench_d3.err; tail -c 200 gpurun_out/r2_bench_d3.log
python tools/op_bench.py --tv --ops roi --rois 512,1024,2048,4096,8192 --json gpurun_out/r2_opbench_roi_608.json > gpurun_out/r2_opbench_roi_608.log 2>&1; tail -1 gpurun_out/r2_opbench_roi_608.log
python tools/op_bench.py --tv --ops roi --height 800 --width 1344 --rois 512,1024,2048,4096,8192 --json gpurun_out/r2_opbench_roi_800.json > gpurun_out/r2_opbench_roi_800.log 2>&1; tail -1 gpurun_out/r2_opbench_roi_800.log
python tools/bwd_check.py --algos 3,4,5 --iters 15 > gpurun_out/r2_bwd_check.log 2>&1; tail -3 gpurun_out/r2_bwd_check.log
ncu --set full --clock-control none --import-source on -k regex:own_bwd -s 2 -c 1 -o gpurun_out/r2_own_bwd -f python tools/bwd_check.py --algos 4 --iters 2 > gpurun_out/r2_ncu_bwd.log 2>&1; tail -1 gpurun_out/r2_ncu_bwd.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1800 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_bench.log 2>&1; wc -l gpurun_out/r2_launches_bench.csv
