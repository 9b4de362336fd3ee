export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_tp.py -x -q -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/c14_tp.txt; cat gpurun_out/c14_tp.txt
