export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -q -s -k "persistent or long_context or split or config" 2>&1 | grep -v "^$" | tail -12 > gpurun_out/c19_pytest.txt; cat gpurun_out/c19_pytest.txt | cut -c1-250
python scripts/ab_variants.py build base: > /dev/null 2>&1
cp qwen3_rs_b200/lib/libqwen3cuda.so qwen3_rs_b200/lib/variant_base.so
AB_REPS=1 timeout 600 python scripts/ab_variants.py run qwen3-8b 32768 24 2>&1 | tail -1
Q3_MEGA_DBG=8 AB_REPS=1 timeout 600 python scripts/ab_variants.py run qwen3-8b 32768 24 2>&1 | tail -1
Q3_MEGA_DBG=9 AB_REPS=1 timeout 600 python scripts/ab_variants.py run qwen3-8b 32768 24 2>&1 | tail -1
