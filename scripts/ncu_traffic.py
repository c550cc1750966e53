"""Turns an `ncu --page raw --csv` dump of ONE generator forward (kernels captured in launch order)
into (1) a per-launch summary CSV and (2) profiles/<tag>_traffic.json, the DRAM bytes per layer that
bench.py reports as `roofline.traffic`.

  ncu -i gpurun_out/x.ncu-rep --page raw --csv > /tmp/raw.csv
  python scripts/ncu_traffic.py /tmp/raw.csv profiles/r01c small 32
"""
import csv
import json
import re
import sys

src, tag, model, batch = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
rows = list(csv.reader(open(src)))
hdr = rows[0]
want = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor']
want = [w for w in want if w in hdr]
idx = [hdr.index(w) for w in want]
units = rows[1]
body = rows[2:]
n_enc = 5 if model == 'small' else 8
layers = (['generator/encoder_%d/conv2d' % i for i in range(1, n_enc + 1)] +
          ['generator/decoder_%d/conv2d_transpose' % k for k in range(n_enc, 0, -1)])
conv = [r for r in body if re.search(r'conv_|deconv_', r[hdr.index('Kernel Name')])]
# the capture window may start mid-forward: rotate so that it begins at encoder_1 (the one-input-channel conv)
conv = conv[-len(layers):]          # one whole forward (the last one captured)
starts = [i for i, r in enumerate(conv) if re.search(r'conv_thin|conv_one_in_tc', r[hdr.index('Kernel Name')])]
if starts:
  conv = conv[starts[0]:] + conv[:starts[0]]
assert len(conv) == len(layers), (len(conv), len(layers))
with open(tag + '_ncu_full_generator_forward.csv', 'w') as f:
  w = csv.writer(f)
  w.writerow(['layer'] + want)
  w.writerow([''] + [units[i] for i in idx])
  for name, r in zip(layers, conv):
    w.writerow([name] + [re.sub(r'advoc::<unnamed>::|void |\(.*', '', r[i]) if j == 0 else r[i]
                         for j, i in enumerate(idx)])
ri, wi = hdr.index('dram__bytes_read.sum'), hdr.index('dram__bytes_write.sum')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
traffic = {}
for name, r in zip(layers, conv):
  traffic[name] = float(r[ri].replace(',', '')) * scale[units[ri]] + float(r[wi].replace(',', '')) * scale[units[wi]]
json.dump({'%s/%d' % (model, batch): traffic}, open(tag + '_traffic.json', 'w'), indent=1)
print('wrote', tag + '_ncu_full_generator_forward.csv', tag + '_traffic.json')
