set -x
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r01_bench.log 2>&1; tail -1 gpurun_out/r01_bench.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --operator pseudo_grid --pseudo-grid-precision bf16 > gpurun_out/r01_bench_pg_bf16.log 2>&1; tail -1 gpurun_out/r01_bench_pg_bf16.log | cut -c1-200
timeout 300 python bench.py --steps 20 --warmup 5 --operator pseudo_grid > gpurun_out/r01_bench_pg_fp32.log 2>&1; tail -1 gpurun_out/r01_bench_pg_fp32.log | cut -c1-200
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_ref.log 2>&1; tail -1 gpurun_out/r01_bench_ref.log | cut -c1-300
