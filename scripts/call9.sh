export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 200 python scripts/diag/prefill_rows.py small 64 3 200 > gpurun_out/c9_rows_new.txt 2>&1; tail -12 gpurun_out/c9_rows_new.txt
Q3_LIB=$PWD/qwen3_rs_b200/lib/variant_pfold.so timeout 200 python scripts/diag/prefill_rows.py small 64 3 200 > gpurun_out/c9_rows_old.txt 2>&1; tail -12 gpurun_out/c9_rows_old.txt
timeout 600 python scripts/ab_prefill.py qwen3-4b 2048 > gpurun_out/c9_ab_prefill.txt 2>&1; cat gpurun_out/c9_ab_prefill.txt
