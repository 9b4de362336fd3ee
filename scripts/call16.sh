export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/c16_pytest_gpu.txt 2>&1; tail -6 gpurun_out/c16_pytest_gpu.txt
( time timeout 900 python bench.py ) > gpurun_out/c16_bench.json 2> gpurun_out/c16_bench.err; tail -4 gpurun_out/c16_bench.err; tail -c 3000 gpurun_out/c16_bench.json
