# GEMM rewrite check: correctness, GEMM-alone timings (new vs old library), whole prefill, decode lazy-sync A/B
set -x
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core_gemm or prefill" 2>&1 | tail -15 > gpurun_out/c1_pytest.txt; cat gpurun_out/c1_pytest.txt
timeout 300 python scripts/bench_gemm.py > gpurun_out/c1_gemm_new.txt 2>&1; cat gpurun_out/c1_gemm_new.txt
Q3_LIB=$PWD/qwen3_rs_b200/lib/variant_pfold.so timeout 300 python scripts/bench_gemm.py > gpurun_out/c1_gemm_old.txt 2>&1; cat gpurun_out/c1_gemm_old.txt
timeout 600 python scripts/ab_prefill.py qwen3-4b 2048 > gpurun_out/c1_ab_prefill.txt 2>&1; cat gpurun_out/c1_ab_prefill.txt
rm -f qwen3_rs_b200/lib/variant_pfold.so
AB_REPS=2 timeout 900 python scripts/ab_variants.py run qwen3-8b 900 64 > gpurun_out/c1_ab_decode.txt 2>&1; cat gpurun_out/c1_ab_decode.txt
