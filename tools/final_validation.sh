# Round-end validation on the GPU box: GPU test suite, smoke, the default bench line (extras included: reference GPU path,
# configs 1 / 3 / 5) and the reference arm.  The ncu launch list is a separate, expensive call (see profiles/README.md).
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -1 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json | cut -c1-200
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; tail -1 gpurun_out/bench_ref.json | cut -c1-200
