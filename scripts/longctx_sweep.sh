# BASELINE config 5: Qwen3-8B decode with a 32K-token KV cache, group-size sweep 32 / 64 / 128
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 700 python scripts/bench_longctx.py qwen3-8b 32768 32,64,128 > gpurun_out/r2_longctx_gs_sweep.jsonl 2> gpurun_out/r2_longctx.err; cat gpurun_out/r2_longctx_gs_sweep.jsonl | cut -c1-400; tail -2 gpurun_out/r2_longctx.err
