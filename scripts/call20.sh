export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests/ -q -s -m gpu ) > gpurun_out/c20_pytest_gpu.txt 2>&1; grep -a "persistent vs graph" gpurun_out/c20_pytest_gpu.txt | cut -c1-200; tail -6 gpurun_out/c20_pytest_gpu.txt
