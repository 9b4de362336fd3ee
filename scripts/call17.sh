export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -k "persistent or long_context or split or layerwise or golden or greedy or config or kv_cache" 2>&1 | tail -6 > gpurun_out/c17_pytest.txt; cat gpurun_out/c17_pytest.txt
AB_REPS=2 timeout 900 python scripts/ab_variants.py run qwen3-8b 900 64 > gpurun_out/c17_ab900.txt 2>&1; cat gpurun_out/c17_ab900.txt
AB_REPS=1 timeout 900 python scripts/ab_variants.py run qwen3-8b 32768 24 > gpurun_out/c17_ab32k.txt 2>&1; cat gpurun_out/c17_ab32k.txt
AB_REPS=1 timeout 900 python scripts/ab_variants.py run qwen3-8b 64 64 > gpurun_out/c17_ab64.txt 2>&1; cat gpurun_out/c17_ab64.txt
