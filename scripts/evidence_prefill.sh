# Round-2 prefill evidence (one B200): ncu full capture of the tcgen05 GEMMs and the tensor-core attention of one Qwen3-4B prefill
# (exported to CSV pages on the box), the GEMM-alone timings / dense ceiling, the TMEM read-bandwidth micro-benchmark.
# scripts/make_profiles_r02.py prefill turns the captures into profiles/r02_ncu_prefill_4b.txt.
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out /tmp/ev
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_q8|k_pf_attention_h" -s 8 -c 5 -f -o /tmp/ev/prefill python scripts/ncu_prefill_target.py qwen3-4b 2048 > gpurun_out/r2_ncu_prefill.log 2>&1; tail -2 gpurun_out/r2_ncu_prefill.log
ncu -i /tmp/ev/prefill.ncu-rep --page raw --csv > gpurun_out/r2_ncu_prefill_raw.csv 2>/dev/null
ncu -i /tmp/ev/prefill.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_ncu_prefill_source.csv.gz
ls -la gpurun_out/r2_ncu_prefill*
timeout 300 python scripts/bench_gemm.py > gpurun_out/r2_gemm_ceilings.txt 2>&1; cat gpurun_out/r2_gemm_ceilings.txt
timeout 120 scripts/micro/tmem_ld_bw > gpurun_out/r2_tmem_ld_bw.txt 2>&1
