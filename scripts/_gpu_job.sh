mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_melspecgan.py -x -q > gpurun_out/s38_pytest.log 2>&1; tail -25 gpurun_out/s38_pytest.log
