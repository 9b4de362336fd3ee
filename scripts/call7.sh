export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python scripts/diag/gemm_trace.py 0 > gpurun_out/c7_trace.txt 2>&1; head -22 gpurun_out/c7_trace.txt | cut -c1-400; tail -1 gpurun_out/c7_trace.txt
