export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
AB_REPS=2 timeout 900 python scripts/ab_variants.py run qwen3-8b 300 64 > gpurun_out/c26_ab300.txt 2>&1; cat gpurun_out/c26_ab300.txt
AB_REPS=2 timeout 900 python scripts/ab_variants.py run qwen3-8b 900 64 > gpurun_out/c26_ab900.txt 2>&1; cat gpurun_out/c26_ab900.txt
