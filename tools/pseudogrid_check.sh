timeout 300 python -m pytest tests/test_gpu_aggregation.py -x -q -m gpu 2>&1 | tail -2
timeout 200 python tools/gpu_probe.py 2>&1 | grep "pseudogrid"
