# Round-2 evidence captures (one B200): sanitizer logs, ncu full capture of the decode kernel (exported to CSV pages here: the
# .ncu-rep files are too large to travel), launch list.  (The prefill captures: scripts/evidence_prefill.sh.)
set -x
export PYTHONUNBUFFERED=1
mkdir -p /tmp/ev gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_target.py micro tiny-untied > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 60 python scripts/sanitize_target.py micro > /tmp/ev/race.log 2>&1; head -c 60000 /tmp/ev/race.log > gpurun_out/r2_sanitizer_racecheck.log; tail -3 /tmp/ev/race.log >> gpurun_out/r2_sanitizer_racecheck.log; grep -c "hazard" /tmp/ev/race.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_target.py micro > gpurun_out/r2_sanitizer_synccheck.log 2>&1; tail -2 gpurun_out/r2_sanitizer_synccheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_mega -s 3 -c 1 -f -o /tmp/ev/mega python scripts/ncu_target.py qwen3-8b 1 6 512 > gpurun_out/r2_ncu_mega.log 2>&1; tail -2 gpurun_out/r2_ncu_mega.log
ncu -i /tmp/ev/mega.ncu-rep --page raw --csv > gpurun_out/r2_ncu_mega_raw.csv 2>/dev/null
ncu -i /tmp/ev/mega.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r2_ncu_mega_source.csv.gz
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --tokens-per-step 16 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 600 python scripts/ab_prefill.py qwen3-4b 2048 2>&1 | head -1 > gpurun_out/r2_prefill_4b.txt; cat gpurun_out/r2_prefill_4b.txt
timeout 600 python scripts/ab_prefill.py qwen3-8b 2048 2>&1 | head -1 > gpurun_out/r2_prefill_8b.txt; cat gpurun_out/r2_prefill_8b.txt
du -sh gpurun_out
