export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/c15_bench_tp2.json 2> gpurun_out/c15_bench_tp2.err; tail -3 gpurun_out/c15_bench_tp2.err; python - <<'PY'
import json
for l in open("gpurun_out/c15_bench_tp2.json"):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "e2e", d["e2e"]["value"], "prefill", d.get("prefill"))
PY
