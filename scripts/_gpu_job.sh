mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_melspecgan.py tests/test_gpu_train.py tests/test_gpu_nets.py -x -q > gpurun_out/s29_pytest.log 2>&1; tail -5 gpurun_out/s29_pytest.log
timeout 300 python bench.py --workload melspecgan --steps 10 > gpurun_out/s29_bench_melspecgan.json 2> gpurun_out/s29_bench_melspecgan.err; tail -3 gpurun_out/s29_bench_melspecgan.err
timeout 300 python bench.py --workload train --model regular --steps 5 > gpurun_out/s29_bench_train_regular.json 2> gpurun_out/s29_bench_train_regular.err
python - <<'PY'
import json
for f in ['s29_bench_melspecgan','s29_bench_train_regular']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
