mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nets.py tests/test_gpu_train.py tests/test_gpu_melspecgan.py -x -q > gpurun_out/s36_pytest.log 2>&1; tail -2 gpurun_out/s36_pytest.log
ADVOC_ONE_PROFILE=1 python scripts/dev_one_prof.py small > gpurun_out/s36_one.txt 2>&1
ADVOC_ONE_PROFILE=1 python scripts/dev_one_prof.py regular >> gpurun_out/s36_one.txt 2>&1
cat gpurun_out/s36_one.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s36_bench.json 2> gpurun_out/s36_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s36_bench.json').read().strip().splitlines()[-1])
r=d.get('roofline') or {}
print(d['value'], d['ms_per_step'])
print('   ', ' '.join('%s=%.0f'%(k.replace('encoder_','e').replace('decoder_','d'),v['us']) for k,v in (r.get('by_layer') or {}).items()))
PY
