export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_cpp_host.py -x -q -k "golden or greedy or forward or chat or generate" 2>&1 | tail -3
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/c25_bench.json 2> gpurun_out/c25_bench.err; python - <<'PY'
import json
for l in open("gpurun_out/c25_bench.json"):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l); print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], "prefill", (d.get("prefill") or {}).get("value"))
PY
