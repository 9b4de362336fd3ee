set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 120 scripts/micro/tmem_ld_bw > gpurun_out/c2_tmem_ld_bw.txt 2>&1; cat gpurun_out/c2_tmem_ld_bw.txt
timeout 200 python scripts/diag/prefill_rows.py > gpurun_out/c2_rows_new.txt 2>&1; tail -5 gpurun_out/c2_rows_new.txt
Q3_LIB=$PWD/qwen3_rs_b200/lib/variant_pfold.so timeout 200 python scripts/diag/prefill_rows.py > gpurun_out/c2_rows_old.txt 2>&1; tail -5 gpurun_out/c2_rows_old.txt
