set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python scripts/bench_gemm.py 0,3,4,2 > gpurun_out/c3_gemm_modes.txt 2>&1; cat gpurun_out/c3_gemm_modes.txt
