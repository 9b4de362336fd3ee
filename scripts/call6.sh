export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core_gemm or prefill" 2>&1 | tail -5 > gpurun_out/c6_pytest.txt; cat gpurun_out/c6_pytest.txt
timeout 300 python scripts/bench_gemm.py 0,1,4,2 > gpurun_out/c6_gemm_modes.txt 2>&1; cat gpurun_out/c6_gemm_modes.txt
