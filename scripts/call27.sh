export PYTHONUNBUFFERED=1
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_tp.py -q -s -k "small8-64-5-8" ) 2>&1 | grep -v "^$" > gpurun_out/r2_tp8_rsag.log; tail -9 gpurun_out/r2_tp8_rsag.log
