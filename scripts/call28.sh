export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/c28_pytest_gpu.txt 2>&1; tail -5 gpurun_out/c28_pytest_gpu.txt
( time timeout 900 python bench.py ) > gpurun_out/c28_bench.json 2> gpurun_out/c28_bench.err; tail -2 gpurun_out/c28_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/c28_bench_ref.json 2> gpurun_out/c28_bench_ref.err; tail -c 600 gpurun_out/c28_bench_ref.json; tail -3 gpurun_out/c28_bench_ref.err
