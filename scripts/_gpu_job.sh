mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s21_pytest.log 2>&1; tail -3 gpurun_out/s21_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/s21_bench.json 2> gpurun_out/s21_bench.err; tail -3 gpurun_out/s21_bench.err
timeout 300 python bench.py --workload train --steps 10 > gpurun_out/s21_bench_train_small.json 2> gpurun_out/s21_bench_train_small.err
timeout 300 python bench.py --workload train --model regular --steps 5 > gpurun_out/s21_bench_train_regular.json 2> gpurun_out/s21_bench_train_regular.err
python - <<'PY'
import json
for f in ['s21_bench','s21_bench_train_small','s21_bench_train_regular']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d['e2e']['value'])
    except Exception as e: print(f, 'ERR', e)
PY
