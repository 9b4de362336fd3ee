mkdir -p gpurun_out
export AB_REPS=1
timeout 500 python scripts/ab_variants.py run qwen3-8b 16 48 2>&1 | grep -v "^\[bench\]" | tee gpurun_out/ab8.txt
timeout 500 python scripts/ab_variants.py run qwen3-8b 480 32 2>&1 | grep -v "^\[bench\]" | tee -a gpurun_out/ab8.txt
