mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s39_pytest.log 2>&1; tail -3 gpurun_out/s39_pytest.log
timeout 300 python bench.py --workload melspecgan --steps 5 > gpurun_out/f_bench_melspecgan_wgangp.json 2> gpurun_out/s39_err.txt; tail -2 gpurun_out/s39_err.txt
python -c "
import json
d=json.loads(open('gpurun_out/f_bench_melspecgan_wgangp.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['launches_per_step'], d['roofline']['achieved'])"
python -c "import __graft_entry__ as g; g.smoke()"
