# Multi-GPU evidence on one box: the tensor-parallel test-suite (tp 2 / 4 / 8 against the oracle, incl. batched prefill under TP) and the
# bench at the box's GPU count.   gpurun --gpus 8 -- 'bash scripts/tp_run.sh'
export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
( time timeout 900 python -m pytest tests/test_gpu_tp.py -q -s ) 2>&1 | grep -v "^$" > gpurun_out/r2_tp_tests_hardware.log; tail -30 gpurun_out/r2_tp_tests_hardware.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2_bench_8b_tp$N.json 2> gpurun_out/r2_bench_tp$N.err; tail -2 gpurun_out/r2_bench_tp$N.err
python - <<PY
import json
for l in open("gpurun_out/r2_bench_8b_tp$N.json"):
    l = l.strip()
    if l.startswith("{"):
        d = json.loads(l); print("N=$N value", d["value"], "e2e", d["e2e"]["value"], "prefill", (d.get("prefill") or {}).get("value"))
PY
