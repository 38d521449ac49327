#!/bin/bash
mkdir -p gpurun_out/r2_bwd_tests
timeout 1200 python -m pytest tests/test_gpu_r2.py tests/test_gpu_sm100.py tests/test_gpu_paged.py -q -m gpu -k "backward or bwd or grad or shim or autograd or golden" > gpurun_out/r2_bwd_tests/pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2_bwd_tests/pytest.log
