set -x
timeout 1200 python -m pytest tests/test_gpu_tp.py -q -s 2>&1 | grep -v "^$" | tail -60 > gpurun_out/r2_tp_tests_hardware.log
tail -25 gpurun_out/r2_tp_tests_hardware.log
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 6 --warmup 3 --no-extras > gpurun_out/r2_bench_tp$n.json 2> gpurun_out/r2_bench_tp$n.err
  tail -2 gpurun_out/r2_bench_tp$n.err
  python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_bench_tp$n.json').read().strip().split('\n')[-1])
print('N=$n value %.1f e2e %.1f us/token %.1f' % (d['value'], d['e2e']['value'], 1e6/d['value']))
"
done
timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline --no-prefill > gpurun_out/r2_bench_tp1.json 2>gpurun_out/r2_bench_tp1.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench_tp1.json').read().strip().split('\n')[-1])
print('N=1 value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))
"
