export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -s -k "prefill" 2>&1 | grep -v "^$" | tail -30 > gpurun_out/c10_pytest.txt; cat gpurun_out/c10_pytest.txt
timeout 600 python scripts/ab_prefill.py qwen3-4b 2048 > gpurun_out/c10_ab_prefill.txt 2>&1; cat gpurun_out/c10_ab_prefill.txt
Q3_PF_ATTN_TF32=1 timeout 600 python scripts/ab_prefill.py qwen3-4b 2048 2>&1 | head -1
