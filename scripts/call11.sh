export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 600 python scripts/ab_prefill.py qwen3-4b 2048 2>&1 | head -1 > gpurun_out/c11_prefill.txt; cat gpurun_out/c11_prefill.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 340 -c 330 --csv --log-file gpurun_out/c11_launches.csv python scripts/ncu_prefill_target.py qwen3-4b 2048 > gpurun_out/c11_ncu.log 2>&1; tail -2 gpurun_out/c11_ncu.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/c11_launches.csv")) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1000.0 if u in ("ns", "nsecond") else v * (1000.0 if u in ("ms", "msecond") else 1.0)
    k = r[ki].split("(")[0][:60]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%4d  %9.1f us  %5.1f %%" % (k, n, t, 100 * t / tot))
print("total %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))
PY
