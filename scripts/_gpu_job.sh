mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_pipeline.py -x -q > gpurun_out/s26_pytest.log 2>&1; tail -5 gpurun_out/s26_pytest.log
timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s26_bench_train_small.json 2> gpurun_out/s26_bench_train_small.err; tail -2 gpurun_out/s26_bench_train_small.err
timeout 300 python bench.py --workload train --model regular --steps 5 > gpurun_out/s26_bench_train_regular.json 2> gpurun_out/s26_bench_train_regular.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s26_launches_train_small.csv python bench.py --workload train --steps 1 --warmup 3 > gpurun_out/s26_l.log 2>&1
python - <<'PY'
import json
for f in ['s26_bench_train_small','s26_bench_train_regular']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
