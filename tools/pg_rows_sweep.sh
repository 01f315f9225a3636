for o in 8 7 6; do echo "dW CTAs/SM=$o"; D3D_PG_DW_CTAS_PER_SM=$o timeout 200 python tools/gpu_probe.py 2>&1 | grep "pseudogrid.*weights tcgen05"; done
timeout 300 python -m pytest tests/test_gpu_aggregation.py -x -q -m gpu 2>&1 | tail -2
