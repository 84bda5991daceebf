# Round-end measurement batch run under gpurun: tests, smoke, both bench arms, op sweep, ncu captures -> gpurun_out/s30_*
set -x
python -m pytest tests/ -m gpu -x -q > gpurun_out/s30_pytest.log 2>&1; tail -3 gpurun_out/s30_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s30_smoke.log 2>&1; tail -1 gpurun_out/s30_smoke.log
python bench.py --steps 3 --warmup 3 > gpurun_out/s30_bench.log 2> gpurun_out/s30_bench.err; tail -c 600 gpurun_out/s30_bench.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/s30_bench_ref.log 2> gpurun_out/s30_bench_ref.err; tail -c 400 gpurun_out/s30_bench_ref.log
python tools/op_bench.py --tv --rois 512,2048 --json gpurun_out/s30_opbench.json > gpurun_out/s30_opbench.log 2>&1; tail -3 gpurun_out/s30_opbench.log
python tools/profile_roi.py > gpurun_out/s30_plain_roi.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:msroi_.*tma -s 4 -c 2 -o gpurun_out/s30_roi -f python tools/profile_roi.py > gpurun_out/s30_ncu_roi.log 2>&1; tail -2 gpurun_out/s30_ncu_roi.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1800 --csv --log-file gpurun_out/s30_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s30_ncu_bench.log 2>&1; tail -2 gpurun_out/s30_ncu_bench.log; wc -l gpurun_out/s30_launches_bench.csv
