mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_nets.py tests/test_gpu_train.py -x -q > gpurun_out/s15_pytest.log 2>&1; tail -3 gpurun_out/s15_pytest.log
ADVOC_P2D_PROFILE=1 python scripts/dev_p2d_prof.py small 32 2>&1 | tail -11 | sed 's/generator.//' | cut -c1-150,330-500 > gpurun_out/s15_prof.txt
cat gpurun_out/s15_prof.txt
timeout 200 python bench.py --no-cpu-baseline > gpurun_out/s15_bench.json 2> gpurun_out/s15_bench.err; tail -3 gpurun_out/s15_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/s15_bench.json').read())
print(d['value'], d['ms_per_step']); print(d['roofline']['by_kernel'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s15_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/s15_l.log 2>&1
