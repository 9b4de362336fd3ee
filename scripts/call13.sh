export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "prefill" 2>&1 | tail -3
timeout 600 python scripts/ab_prefill.py qwen3-4b 2048 2>&1 | head -1
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_pf_attention_h -s 40 -c 2 python scripts/ncu_prefill_target.py qwen3-4b 2048 2>&1 | grep -E "gpu__time|tensor_cycles" 
