export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 300 python scripts/profile_mega.py qwen3-8b 900 > gpurun_out/r2_phase_pos900_final.txt 2>&1; head -40 gpurun_out/r2_phase_pos900_final.txt
timeout 300 python scripts/profile_mega.py qwen3-8b 64 > gpurun_out/r2_phase_pos64_final.txt 2>&1; grep "per layer" gpurun_out/r2_phase_pos64_final.txt
