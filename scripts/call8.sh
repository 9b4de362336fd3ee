export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
Q3_LIB=$PWD/qwen3_rs_b200/lib/variant_noscale.so timeout 300 python scripts/bench_gemm.py 0,4,2 > gpurun_out/c8_noscale.txt 2>&1; cat gpurun_out/c8_noscale.txt
