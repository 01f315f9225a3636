# Round-end validation on the GPU box: GPU test suite, smoke, the bench lines kept under profiles/, the launch list.
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r01_bench.log 2>&1; tail -1 gpurun_out/r01_bench.log | cut -c1-160
timeout 300 python bench.py --steps 20 --warmup 5 --operator pseudo_grid --pseudo-grid-precision bf16 > gpurun_out/r01_bench_pg_bf16.log 2>&1; tail -1 gpurun_out/r01_bench_pg_bf16.log | cut -c1-160
timeout 300 python bench.py --steps 20 --warmup 5 --operator pseudo_grid > gpurun_out/r01_bench_pg_fp32.log 2>&1; tail -1 gpurun_out/r01_bench_pg_fp32.log | cut -c1-160
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_ref.log 2>&1; tail -1 gpurun_out/r01_bench_ref.log | cut -c1-160
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_r01_v5.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; wc -l gpurun_out/launches_r01_v5.csv
