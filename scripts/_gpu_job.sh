mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s18_pytest.log 2>&1; tail -3 gpurun_out/s18_pytest.log
timeout 300 python bench.py > gpurun_out/s18_bench.json 2> gpurun_out/s18_bench.err; tail -3 gpurun_out/s18_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s18_bench_n2.json 2> gpurun_out/s18_bench_n2.err; tail -3 gpurun_out/s18_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 5 --warmup 3 > gpurun_out/s18_bench_train_n2.json 2> gpurun_out/s18_bench_train_n2.err; tail -3 gpurun_out/s18_bench_train_n2.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s18_bench_ref.json 2> gpurun_out/s18_bench_ref.err
python - <<'PY'
import json
for f in ['s18_bench','s18_bench_n2','s18_bench_train_n2','s18_bench_ref']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
        r=d.get('roofline') or {}
        print(f, d['value'], d['ms_per_step'], d.get('e2e'), {k:v for k,v in r.items() if k not in ('by_kernel','by_layer')})
        for k,v in (r.get('by_layer') or {}).items(): print('   ',k,v)
    except Exception as e: print(f, 'ERR', e)
PY
