mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s33_pytest.log 2>&1; tail -2 gpurun_out/s33_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s33_bench.json 2> gpurun_out/s33_bench.err
timeout 300 python bench.py --no-cpu-baseline --model regular --steps 20 > gpurun_out/s33_bench_reg.json 2> gpurun_out/s33_bench_reg.err
timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s33_bench_train_small.json 2> gpurun_out/s33_bench_train_small.err
timeout 300 python bench.py --workload train --model regular --steps 5 > gpurun_out/s33_bench_train_regular.json 2> gpurun_out/s33_bench_train_regular.err
python - <<'PY'
import json
for f in ['s33_bench','s33_bench_reg','s33_bench_train_small','s33_bench_train_regular']:
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    r=d.get('roofline') or {}
    print(f, d['value'], d['ms_per_step'])
    print('   ', ' '.join('%s=%.0f'%(k.replace('encoder_','e').replace('decoder_','d'),v['us']) for k,v in (r.get('by_layer') or {}).items()))
PY
