export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -k "persistent or long_context or split or config or fast_mode" 2>&1 | tail -3
AB_REPS=1 timeout 900 python scripts/ab_variants.py run qwen3-8b 32768 24 > gpurun_out/c22_ab32k.txt 2>&1; cat gpurun_out/c22_ab32k.txt
AB_REPS=2 timeout 900 python scripts/ab_variants.py run qwen3-8b 900 64 > gpurun_out/c22_ab900.txt 2>&1; cat gpurun_out/c22_ab900.txt
AB_REPS=1 timeout 900 python scripts/ab_variants.py run qwen3-8b 4000 32 > gpurun_out/c22_ab4k.txt 2>&1; cat gpurun_out/c22_ab4k.txt
