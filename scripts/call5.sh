export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python scripts/diag/gemm_trace.py 0,4,5 > gpurun_out/c5_trace.txt 2>&1; cat gpurun_out/c5_trace.txt
