# bench lines of the smaller BASELINE shapes (configs 2 and 3) with the final kernels
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --model qwen3-0.6b --no-extras --steps 6 > gpurun_out/r2_bench_06b.json 2> gpurun_out/r2_bench_06b.err; tail -1 gpurun_out/r2_bench_06b.err
timeout 300 python bench.py --model qwen3-4b --no-extras --steps 6 > gpurun_out/r2_bench_4b.json 2> gpurun_out/r2_bench_4b.err; tail -1 gpurun_out/r2_bench_4b.err
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_06b.json", "gpurun_out/r2_bench_4b.json"):
    for l in open(f):
        l = l.strip()
        if l.startswith("{"):
            d = json.loads(l); print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 3), "prefill", round((d.get("prefill") or {}).get("value", 0)), "cpu", (d.get("cpu_baseline") or {}).get("value"), "parity fast", (d.get("parity") or {}).get("fast", {}).get("max_abs_dlogit"))
PY
