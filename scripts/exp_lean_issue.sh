#!/bin/bash
# GPU experiment queued at the end of round 1 (the budget ran out before it could run):
#   /usr/local/graft/bin/gpurun --timeout 600 -- 'bash scripts/exp_lean_issue.sh'
# 1. parity of the patch kernel's lean issue loop (ADVOC_P2D_LEAN_ISSUE=1; DESIGN.md section 7,
#    profiles/r01c_mma_issue_stall_samples.txt), 2. forward / train bench with and without it,
# 3. the same with every eligible layer forced onto the patch kernel (ADVOC_P2D_FORCE=1: lifts the
#    "filter bytes per tile <= 512 KB" rule that keeps decoder_3/4 and encoder_4/5 on the per-tap kernel).
mkdir -p gpurun_out
ADVOC_P2D_LEAN_ISSUE=1 timeout 400 python -m pytest tests/test_gpu_nets.py tests/test_gpu_forced_paths.py \
  tests/test_gpu_train.py tests/test_gpu_melspecgan.py -x -q > gpurun_out/lean_pytest.log 2>&1
tail -3 gpurun_out/lean_pytest.log
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  export ADVOC_P2D_LEAN_ISSUE=$1
  if [ "$2" = 1 ]; then export ADVOC_P2D_FORCE=1; else unset ADVOC_P2D_FORCE; fi
  for w in infer train; do
    timeout 200 python bench.py --workload $w --no-cpu-baseline > gpurun_out/lean_${w}_$1$2.json 2> gpurun_out/lean_${w}_$1$2.err
    python - "$w" "$1" "$2" <<'PY'
import json, sys
w, lean, force = sys.argv[1:4]
try:
  d = json.loads(open('gpurun_out/lean_%s_%s%s.json' % (w, lean, force)).read().strip().splitlines()[-1])
  layers = d.get('roofline', {}).get('by_layer', {})
  print('lean=%s force=%s %-5s value %.4g  ms/step %.3f' % (lean, force, w, d['value'], d['ms_per_step']),
        ' '.join('%s:%s/%.0f' % (k.replace('coder_', ''), v['kernel'][:7], v['us']) for k, v in layers.items()))
except Exception as e:
  print('lean=%s force=%s %s failed: %r' % (lean, force, w, e))
PY
  done
done | tee gpurun_out/lean_summary.txt
# 4. finer than FORCE: sweep the filter-bytes threshold of the routing rule with the lean loop on
export ADVOC_P2D_LEAN_ISSUE=1
unset ADVOC_P2D_FORCE
for kb in 1024 2048 4096; do
  ADVOC_P2D_MAX_FILTER_KB=$kb timeout 200 python bench.py --no-cpu-baseline > gpurun_out/lean_infer_kb$kb.json 2> gpurun_out/lean_infer_kb$kb.err
  python -c "
import json
d=json.loads(open('gpurun_out/lean_infer_kb$kb.json').read().strip().splitlines()[-1]); print('max_filter_kb=$kb', d['value'], d['ms_per_step'])" | tee -a gpurun_out/lean_summary.txt
done

# 5. deep layers on the patch kernel with narrower N tiles (more tiles, double-buffered accumulators)
for bn in 64 128; do
  ADVOC_P2D_FORCE=1 ADVOC_P2D_MAX_BN=$bn timeout 200 python bench.py --no-cpu-baseline > gpurun_out/lean_infer_bn$bn.json 2> gpurun_out/lean_infer_bn$bn.err
  python -c "
import json
d=json.loads(open('gpurun_out/lean_infer_bn$bn.json').read().strip().splitlines()[-1]); print('force, max_bn=$bn', d['value'], d['ms_per_step'], {k: round(v['us']) for k, v in d['roofline']['by_layer'].items()})" | tee -a gpurun_out/lean_summary.txt
done
