#!/usr/bin/env python
"""Headline benchmark: AdVoc generator forward, mel-frames/sec (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            # this framework (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the TF1 graph

One step = one pass of the hot path over one batch of synthetic input: linear mel
[B=32, 256, 80] -> pinv lift -> AdVoc-small U-Net -> magnitude [32, 256, 513]; dropout on
decoder_5/4 active as in the reference (advoc_model_small.py:149-154).  `value` is measured with
the mel batch already resident in HBM (CUDA events, L2 flushed between steps); `e2e` is the same
metric through the public host API (`advoc_b200.infer.MelToMag.__call__`) with pinned host
buffers, H2D and D2H copies inside the timed region.  N > 1 runs one replica per GPU (weak
scaling, no data-path collective: inference chunks are independent).  `--workload train`
times the G+D train step instead (BASELINE configs[2]/[3]).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

FLOP_PER_FRAME = {'small': 21.84e6, 'regular': 95.75e6}   # SURVEY.md section 8(d), G forward
T = 256


def _peaks():
  try:
    with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
      p = json.load(f)
    return dict(hbm=p['hbm_gbs'], tensor=p['bf16_tflops'], tensor_sustained=p['bf16_tflops_sustained'],
                source='measured')
  except Exception:
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source='fallback')


class ClockSampler(object):
  """Samples SM clock + throttle reasons via NVML while the timed region runs."""
  BAD = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown'}
  NOTE = {0x4: 'sw_power_cap'}

  def __init__(self, index):
    self.samples, self.reasons, self.max_mhz = [], set(), None
    self._stop = threading.Event()
    self._t = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:
      self.nv = None

  def _run(self):
    nv = self.nv
    while not self._stop.is_set():
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in list(self.BAD.items()) + list(self.NOTE.items()):
          if r & bit:
            self.reasons.add(name)
      except Exception:
        pass
      time.sleep(0.005)

  def __enter__(self):
    if self.nv is not None:
      self._t = threading.Thread(target=self._run, daemon=True)
      self._t.start()
    return self

  def __exit__(self, *a):
    self._stop.set()
    if self._t is not None:
      self._t.join()

  def summary(self):
    s = sorted(self.samples)
    return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz,
                reasons=sorted(self.reasons), samples=len(s))


def _model(name):
  from advoc_b200.model import Advoc, AdvocSmall, Modes
  m = AdvocSmall(Modes.INFER) if name == 'small' else Advoc(Modes.INFER)
  m.init_params(seed=0)
  return m


def _synthetic_mel(batch, seed=1):
  import torch
  g = torch.Generator().manual_seed(seed)
  return torch.randn(batch, T, 80, generator=g).abs()


# ---------------------------------------------------------------------------
# CPU baseline (oracle port of the TF1 graph; bench.py's cpu_baseline leg and --impl reference)
# ---------------------------------------------------------------------------
def cpu_generator_frames_per_sec(model_name, sample_batch, repeats, warmup=1):
  import numpy as np
  import torch
  from oracle import nets_torch as O
  from oracle import spectral_np as OS
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  spec = O.SMALL if model_name == 'small' else O.REGULAR
  P = O.init_params(spec, seed=0)
  Winv = torch.from_numpy(
      OS.create_inverse_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80)
      .astype(np.float32))
  mel = _synthetic_mel(sample_batch)
  g = torch.Generator().manual_seed(7)
  shapes = {}
  with torch.no_grad():
    x = (mel @ Winv.t()).unsqueeze(-1)
    _, layers = O.generator(P, x[:1], spec, return_layers=True)
    n_enc = spec.num_enc_layers + 1
    for j, k in enumerate(range(n_enc, 1, -1)):
      if k in spec.dropout_decoders:
        shapes[k] = (sample_batch,) + tuple(layers[n_enc + j].shape[1:])
  times = []
  with torch.no_grad():
    for it in range(warmup + repeats):
      t0 = time.perf_counter()
      x = (mel @ Winv.t()).unsqueeze(-1)
      masks = {k: (torch.rand(s, generator=g) < 0.5).float() for k, s in shapes.items()}
      O.generator(P, x, spec, masks)
      dt = time.perf_counter() - t0
      if it >= warmup:
        times.append(dt)
  return sample_batch * T / min(times), sum(times) / len(times), cores


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  sample = 8 if args.model == 'small' else 2
  fps, mean_s, cores = cpu_generator_frames_per_sec(args.model, sample, max(args.steps, 1),
                                                    max(args.warmup, 1))
  desc = ('PyTorch-CPU fp32 restatement of the TF1 graph (oracle/nets_torch.py; TF1/lws/librosa '
          'not installable), %d of %d samples per step' % (sample, args.batch))
  line = {
      'impl': 'reference', 'metric': 'mel-frames/sec generator fwd', 'value': fps,
      'unit': 'mel-frames/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': mean_s * 1e3 * args.batch / sample, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': _config(args),
      'cpu_baseline': {'value': fps, 'unit': 'mel-frames/s', 'cores': cores, 'kind': 'port',
                       'sample': desc},
      'e2e': {'value': fps, 'unit': 'mel-frames/s', 'h2d_bytes_per_step': 0,
              'd2h_bytes_per_step': 0},
      'gpu_launches': 0,
  }
  print(json.dumps(line))


def _config(args):
  return {'workload': 'AdVoc-%s generator forward, synthetic |N(0,1)| linear mel [%d, 80, 256] -> '
                      'magnitude [%d, 513, 256], random N(0,0.02) weights, dropout on'
                      % (args.model, args.batch, args.batch),
          'batch_per_gpu': args.batch, 'frames_per_sample': T,
          'l2': 'flushed between timed steps (256 MiB write); per-step working set %d MB > 126 MB L2'
                % (17 * args.batch if args.model == 'small' else 33 * args.batch)}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_gpu(args):
  import torch
  import torch.distributed as dist
  from advoc_b200 import _native as N
  from advoc_b200.infer import MelToMag

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise RuntimeError('bench.py needs a GPU (the product path has no CPU fallback)')
  torch.cuda.set_device(local)
  bound = 0
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    if not os.environ.get('ADVOC_NO_AFFINITY'):
      # one process per GPU: keep each rank's pinned buffers and copy threads on its GPU's NUMA node
      from advoc_b200 import dist as D
      bound = D.bind_to_gpu_numa(local)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  model = _model(args.model)
  math = {'f16': N.MATH_F16, 'tf32': N.MATH_AUTO, 'fp32': N.MATH_FP32}[args.math]
  eng = MelToMag(model, args.batch, 'linear', dropout='rng', use_graph=not args.no_graph, math=math)
  mel_h = _synthetic_mel(args.batch, seed=1 + rank).pin_memory()
  eng.mel_d.copy_(mel_h)
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  stream = torch.cuda.current_stream()

  def timed(fn, steps, warmup):
    for _ in range(warmup):
      fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(steps)]
    barrier()
    for a, b in evs:
      flush.zero_()
      a.record(stream)
      fn()
      b.record(stream)
    barrier()
    ms = [a.elapsed_time(b) for a, b in evs]
    return sum(ms)

  n0 = N.launch_count()
  eng.step_device()
  per_step = eng.launches_per_step or (N.launch_count() - n0)

  def e2e_stream(steps, warmup):
    """`steps` batches through the public streaming API (host buffers in, host buffers out; the
    copies of neighbouring batches overlap the forward).  Every batch pays its own H2D and D2H
    inside the timed region; timed on the device between the first copy and the last result."""
    for _ in eng.run_stream([mel_h] * max(warmup, 2)):
      pass
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    n = 0
    for out in eng.run_stream([mel_h] * steps):
      n += 1
    b.record(stream)
    barrier()
    assert n == steps
    return a.elapsed_time(b)

  with ClockSampler(local) as clk:
    total_ms = timed(eng.step_device, args.steps, args.warmup)
    e2e_serial_ms = timed(lambda: eng(mel_h), args.steps, args.warmup)
    e2e_ms = e2e_stream(args.steps, args.warmup)
  if world > 1:
    t = torch.tensor([total_ms, e2e_ms, e2e_serial_ms], device='cuda', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, e2e_serial_ms = float(t[0]), float(t[1]), float(t[2])

  frames = world * args.batch * T * args.steps
  value = frames / (total_ms * 1e-3)
  e2e = frames / (e2e_ms * 1e-3)

  # per-layer device time of one un-graphed step (CUDA events on the launching stream):
  # the roofline is reported for the dominant kernel
  roof = layer_roofline(eng, args, stream, flush)
  parity = infer_parity(eng, model, args, mel_h)
  flags = N.debug_flags()
  if flags:
    raise RuntimeError('a tcgen05 pipeline wait timed out during the benchmark (code %d): numbers invalid' % flags)

  line = {
      'metric': 'mel-frames/sec generator fwd', 'value': value, 'unit': 'mel-frames/s',
      'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
      'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None,
      'dtype': {'f16': 'f16 operands / storage between tensor-core layers, fp32 accumulate (tcgen05 kind::f16)',
                'tf32': 'tf32 (fp32 accumulate, fp32 storage)', 'fp32': 'f32 (CUDA cores)'}[args.math],
      'data': 'synthetic',
      'config': _config(args),
      'e2e': {'value': e2e, 'unit': 'mel-frames/s', 'ms_per_step': e2e_ms / args.steps,
              'h2d_bytes_per_step': int(mel_h.numel() * 4),
              'd2h_bytes_per_step': int(eng.out_h.numel() * 4),
              'api': 'infer.MelToMag.run_stream (pinned host in / out, copies overlapped with the '
                     'neighbouring batches)',
              'cpu_cores_bound_per_rank': bound,
              'serial_call_value': frames / (e2e_serial_ms * 1e-3),
              'serial_call_api': 'infer.MelToMag.__call__ (H2D, forward, D2H back to back)'},
      'gpu_launches': int(per_step * args.steps * 2),
      'launches_per_step': int(per_step),
      'clocks': clk.summary(),
      'roofline': roof,
      'step_tflops': FLOP_PER_FRAME[args.model] * frames / (total_ms * 1e-3) / 1e12,
      'parity_rel_l2': parity,
  }
  if not args.no_train_step:
    # the multi-GPU object of this framework is the data-parallel train step (BASELINE configs[2]/[3]):
    # AdVoc-regular, 32 samples per GPU, NCCL gradient all-reduce inside the timed region.  `value`
    # above stays the generator forward (replicas), so the N = 1 line still matches BENCH.
    line['train_step'] = measure_train_step('regular', 32, world, rank, 10, 3, barrier, with_e2e=False)
  if rank == 0:
    if not args.no_cpu_baseline and world == 1:
      sample = 8 if args.model == 'small' else 2
      fps, mean_s, cores = cpu_generator_frames_per_sec(args.model, sample, 3, 1)
      line['cpu_baseline'] = {
          'value': fps, 'unit': 'mel-frames/s', 'cores': cores, 'kind': 'port',
          'sample': 'oracle/nets_torch.py (PyTorch-CPU fp32 port of the TF1 graph), %d of %d '
                    'samples, best of 3' % (sample, args.batch)}
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def infer_parity(eng, model, args, mel_h, samples=2):
  """GPU vs oracle on `samples` samples of the batch bench.py times, dropout off (the oracle cannot
  reproduce the counter-based masks): relative L2 of the generated magnitudes, through the very
  engine (same kernels, same tile shapes, same batch size) that was timed."""
  import numpy as np
  import torch
  from oracle import nets_torch as O
  from oracle import spectral_np as OS
  spec = O.SMALL if args.model == 'small' else O.REGULAR
  P = {k: v.detach().cpu() for k, v in model.params.items()}
  keep = eng.dropout
  eng.dropout = None
  try:
    eng.mel_d.copy_(mel_h)
    got = eng._launch(1)[:samples, :, :, 0].float().cpu()
  finally:
    eng.dropout = keep
  Winv = torch.from_numpy(OS.create_inverse_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80)
                          .astype(np.float32))
  with torch.no_grad():
    x = (mel_h[:samples] @ Winv.t()).unsqueeze(-1)
    ref = O.generator(P, x, spec)[..., 0]
  return float((got.double() - ref.double()).norm() / ref.double().norm())


def _tf32_peak():
  """Dense TF32 tensor-core peak measured live the way MEASURED_PEAKS.json measures bf16
  (cuBLAS 8192^3, best of 10, CUDA events): the conv kernels compute in kind::tf32, so this is
  the roofline denominator BASELINE.md section 2 asks for."""
  import torch
  old = torch.backends.cuda.matmul.allow_tf32
  torch.backends.cuda.matmul.allow_tf32 = True
  try:
    n = 8192
    a = torch.randn(n, n, device='cuda')
    b = torch.randn(n, n, device='cuda')
    c = torch.empty(n, n, device='cuda')
    best = float('inf')
    for i in range(12):
      s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      s.record()
      torch.matmul(a, b, out=c)
      e.record()
      torch.cuda.synchronize()
      if i >= 2:
        best = min(best, s.elapsed_time(e))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12
  finally:
    torch.backends.cuda.matmul.allow_tf32 = old


def _ncu_issue_bound(kernel):
  """For a kernel that neither HBM nor the tensor pipe bounds: the busiest SM resource from the latest
  committed `ncu --set full` summary (profiles/*_ncu_full_*.csv), as {bound, frac, source}."""
  import csv
  import glob
  best = None
  for fn in sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_ncu_full_*.csv'))):
    try:
      rows = list(csv.reader(open(fn)))
      hdr = rows[0]
      ki, ii = hdr.index('Kernel Name'), hdr.index('sm__issue_active.avg.pct_of_peak_sustained_elapsed')
      vals = [float(r[ii]) for r in rows[2:] if kernel in r[ki]]
      if vals:
        best = {'bound': 'instruction issue (sm__issue_active, % of peak)', 'frac': max(vals) / 100.0,
                'source': os.path.basename(fn)}
    except Exception:
      pass
  return best


def _ncu_traffic(model, batch):
  """DRAM bytes per launch from the committed `ncu --set full` capture of this workload
  (profiles/*_traffic.json: {"model/batch": {layer name: dram read+write bytes}})."""
  import glob
  best = {}
  for fn in sorted(glob.glob(os.path.join(ROOT, 'profiles', '*_traffic.json'))):
    try:
      with open(fn) as f:
        best = json.load(f).get('%s/%d' % (model, batch), best)
    except Exception:
      pass
  return best


def layer_roofline(eng, args, stream, flush):
  """Times every layer launch of one step with CUDA events on the launching stream (L2 flushed
  before each step) and reports the roofline of the kernel with the largest share of the step."""
  import torch
  G = eng.G
  recs = []
  orig = G._run_layer

  def timed_run(L, *a):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(stream)
    orig(L, *a)
    e.record(stream)
    recs.append((L, s, e))

  peaks = _peaks()
  acc, per_layer = {}, {}
  reps = 5
  G._run_layer = timed_run
  try:
    for _ in range(reps):
      flush.zero_()
      eng._launch(1)
    torch.cuda.synchronize()
  finally:
    G._run_layer = orig
  for L, s, e in recs:
    k = L.kernel_family().split('/')[0]
    inst = k + ('<%d,%s>' % (L.tile_n(), 'f16' if L.half_operands() else 'tf32') if L.tile_n() else '')
    ms = s.elapsed_time(e) / reps
    d = acc.setdefault(k, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0.0, layers=[]))
    d['ms'] += ms
    d['flops'] += L.flops() / reps
    d['bytes'] += L.algorithmic_bytes() / reps
    d['launches'] += 1.0 / reps
    if L.name not in d['layers']:
      d['layers'].append(L.name)
    pl = per_layer.setdefault(L.name, dict(kernel=inst, ms=0.0, gflop=L.flops() / 1e9,
                                           mbytes=L.algorithmic_bytes() / 1e6))
    pl['ms'] += ms
  total = sum(d['ms'] for d in acc.values())
  top = max(acc, key=lambda k: acc[k]['ms'])
  d = acc[top]
  kernels = {k: {'ms': round(v['ms'], 4), 'share': round(v['ms'] / total, 3),
                 'launches_per_step': int(round(v['launches'])),
                 'tflops': round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 2),
                 'gbs': round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 1)} for k, v in acc.items()}
  layers = {n.replace('generator/', '').split('/')[0]:
                {'kernel': v['kernel'], 'us': round(v['ms'] * 1e3, 1),
                 'tflops': round(v['gflop'] / v['ms'], 2), 'gbs': round(v['mbytes'] / v['ms'], 1)}
            for n, v in per_layer.items()}
  traffic_by_layer = _ncu_traffic(args.model, args.batch)
  traffic = [traffic_by_layer[n] for n in d['layers'] if n in traffic_by_layer]
  traffic = (sum(traffic) / len(traffic)) if len(traffic) == len(d['layers']) and traffic else None
  n_launch = max(d['launches'], 1.0)
  if (top.startswith('conv_p2d') or top.startswith('conv_tc')) and args.math == 'f16':
    tf32 = _tf32_peak()
    ach = d['flops'] / (d['ms'] * 1e-3) / 1e12
    return {'kernel': top, 'bound': 'tensor', 'achieved': ach, 'peak': peaks['tensor'], 'unit': 'TFLOP/s',
            'frac': ach / peaks['tensor'], 'traffic': traffic,
            'peak_source': 'MEASURED_PEAKS.json bf16 burst (%s): the kernel computes in kind::f16, whose dense '
                           'rate equals bf16\'s; cuBLAS TF32 8192^3 measured in this run = %.1f'
                           % (peaks['source'], tf32),
            'frac_of_tf32_peak': ach / tf32,
            'flops_per_launch': d['flops'] / n_launch, 'ms_per_launch': d['ms'] / n_launch,
            'share_of_step': d['ms'] / total, 'by_kernel': kernels, 'by_layer': layers}
  if top.startswith('conv_p2d') or top.startswith('conv_tc'):
    tf32 = _tf32_peak()
    ach = d['flops'] / (d['ms'] * 1e-3) / 1e12
    return {'kernel': top, 'bound': 'tensor', 'achieved': ach, 'peak': tf32, 'unit': 'TFLOP/s',
            'frac': ach / tf32, 'traffic': traffic,
            'peak_source': 'cuBLAS TF32 8192^3 measured in this run (the kernel computes in '
                           'kind::tf32); MEASURED_PEAKS.json bf16 burst = %.1f' % peaks['tensor'],
            'frac_of_bf16_peak': ach / peaks['tensor'],
            'note': ('the per-tap kernel (conv_tc_kernel) fetches every input pixel once per filter tap through L2; '
                     'with fp32 operands its wide layers run at 81-100 % of the ~6300 B/clk L2 throughput cap '
                     '(profiles/r01c_l2_delivery.txt)') if top.startswith('conv_tc') else None,
            'flops_per_launch': d['flops'] / n_launch, 'ms_per_launch': d['ms'] / n_launch,
            'share_of_step': d['ms'] / total, 'by_kernel': kernels, 'by_layer': layers}
  ach = d['bytes'] / (d['ms'] * 1e-3) / 1e9
  return {'kernel': top, 'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm'], 'unit': 'GB/s',
          'frac': ach / peaks['hbm'], 'traffic': traffic, 'peak_source': peaks['source'] + ' copy bandwidth',
          'bytes_per_launch': d['bytes'] / n_launch, 'ms_per_launch': d['ms'] / n_launch,
          'share_of_step': d['ms'] / total, 'by_kernel': kernels, 'by_layer': layers}


# ---------------------------------------------------------------------------
# train-step workload (BASELINE configs[2]/[3]): 1 step = D update on batch a + G update on
# batch b (advoc_model.py:285-289); per sample 4*G_f + 8*D_f FLOPs
# ---------------------------------------------------------------------------
TRAIN_GFLOP_PER_SAMPLE = {'small': 48.0, 'regular': 199.3}


def _train_config(args, world):
  return {'workload': 'AdVoc-%s G+D train step (log-GAN + 10*L1, TF1 Adam), synthetic magnitude '
                      'batches [%d, 256, 513] per GPU, two minibatches per step (D then G)'
                      % (args.model, args.batch),
          'batch_per_gpu': args.batch, 'global_batch': args.batch * world,
          'l2': 'per-step working set > 1 GB >> 126 MB L2, no explicit flush',
          'parallelism': 'dp%d' % world}


def _synthetic_mag(batch, seed):
  import torch
  g = torch.Generator().manual_seed(seed)
  return (torch.randn(batch, T, 513, 1, generator=g).abs() * 0.1)


def _train_engine(model_name, batch, world, rank, overlap=True, use_graphs=True):
  from advoc_b200 import nets
  from advoc_b200.train import TrainEngine
  small = model_name == 'small'
  spec = nets.GenSpec(32 if small else 64, 5 if small else 8, (5, 4) if small else (8, 7, 6))
  ndf = 32 if small else 64
  params = nets.init_params(spec.ngf, ndf, spec.n_enc, seed=0)
  return TrainEngine(spec, ndf, params, batch, world_size=world, rank=rank, overlap=overlap, use_graphs=use_graphs)


def measure_train_step(model_name, batch, world, rank, steps, warmup, barrier, with_e2e=True, overlap=True,
                       use_graphs=True):
  """Times the data-parallel G+D train step (advoc_model.py:285-289: D update on one minibatch, G update
  on the next), `batch` samples per GPU, the NCCL gradient all-reduces inside the timed region.  Returns
  a dict: samples/s over all ranks (max-over-ranks time), ms/step, the all-reduce time the compute
  stream was left waiting for (events around every wait), launches per step."""
  import torch
  import torch.distributed as dist
  from advoc_b200 import _native as N
  from advoc_b200.model import SpectralUtil
  eng = _train_engine(model_name, batch, world, rank, overlap, use_graphs)
  su = SpectralUtil()
  mags_h = [_synthetic_mag(batch, 100 + 2 * rank + i).pin_memory() for i in range(2)]
  mags_d = [m.cuda() for m in mags_h]
  stage = [torch.empty_like(m) for m in mags_d]
  stream = torch.cuda.current_stream()

  def lift(mag):
    return su.mel_linear_to_mag_spec(su.mag_to_mel_linear_spec(mag))

  def step_device():
    xa, xb = lift(mags_d[0]), lift(mags_d[1])
    eng.train_loop((xa, mags_d[0]), (xb, mags_d[1]))

  def step_e2e():
    for i in range(2):
      stage[i].copy_(mags_h[i], non_blocking=True)
    xa, xb = lift(stage[0]), lift(stage[1])
    eng.train_loop((xa, stage[0]), (xb, stage[1]))
    return eng.loss_values()          # D2H read of the step's losses

  def timed(fn, n, w):
    for _ in range(w):
      fn()
    eng._finish_d()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n):
      fn()
    eng._finish_d()
    b.record(stream)
    barrier()
    return a.elapsed_time(b)

  # kernels of one steady-state step: launches the library counted directly (plus, for engines that
  # replay CUDA graphs, the recorded launch counts per graph)
  step_device()
  torch.cuda.synchronize()
  n0, r0, c0 = N.launch_count(), getattr(eng, 'replayed_launches', 0), getattr(eng, 'capture_launches', 0)
  step_device()
  torch.cuda.synchronize()
  per_step = ((N.launch_count() - n0) - (getattr(eng, 'capture_launches', 0) - c0)
              + (getattr(eng, 'replayed_launches', 0) - r0))
  total_ms = timed(step_device, steps, warmup)
  eng.collective_events = []
  exposed_run_ms = timed(step_device, steps, 0)
  torch.cuda.synchronize()
  exposed_ms = sum(a.elapsed_time(b) for a, b in eng.collective_events) / steps
  eng.collective_events = None
  e2e_ms = timed(step_e2e, steps, warmup) if with_e2e else None
  # the same collectives alone, back to back (no compute to hide behind): their isolated cost
  iso_ms = 0.0
  if world > 1:
    lo_d, hi_d = eng.flat.dis_range()
    lo_g, hi_g = eng.flat.gen_range()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(5):
      dist.all_reduce(eng.flat.g[lo_d:hi_d])
      dist.all_reduce(eng.flat.g[lo_g:hi_g])
    b.record(stream)
    barrier()
    iso_ms = a.elapsed_time(b) / 5
  vals = [total_ms, exposed_ms, iso_ms, e2e_ms if e2e_ms is not None else 0.0, exposed_run_ms]
  if world > 1:
    t = torch.tensor(vals, device='cuda', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    vals = [float(v) for v in t]
  total_ms, exposed_ms, iso_ms, e2e_max, exposed_run_ms = vals
  flags = N.debug_flags()
  if flags:
    raise RuntimeError('a tcgen05 pipeline wait timed out during the train benchmark (code %d)' % flags)
  samples = world * batch * steps
  value = samples / (total_ms * 1e-3)
  tflops = TRAIN_GFLOP_PER_SAMPLE[model_name] * 1e9 * value / 1e12
  out = {'model': 'AdVoc-' + model_name, 'batch_per_gpu': batch, 'global_batch': batch * world, 'n_gpus': world,
         'value': value, 'unit': 'samples/s', 'ms_per_step': total_ms / steps, 'steps': steps,
         'allreduce_bytes_per_step': int(4 * eng.flat.total),
         'allreduce_exposed_ms_per_step': exposed_ms, 'allreduce_isolated_ms_per_step': iso_ms,
         'allreduce': ('NCCL sum-all-reduce of the flat fp32 gradient buffer: D slice in flight under the next generator '
                       'forward, G slice in three buckets (decoders | encoder_n..5 | encoder_4..1) behind the encoder '
                       'backward; exposed = time the compute stream waited') if (world > 1 and overlap) else
                      ('NCCL, not overlapped' if world > 1 else 'none (1 GPU)'),
         'tflops_per_gpu': tflops / world, 'launches_per_step': int(per_step),
         'launch': 'CUDA-graph replay of the segments between collectives' if use_graphs else 'eager',
         'loss_values': [float(v) for v in eng.loss_values()]}
  if with_e2e:
    out['e2e'] = {'value': samples / (e2e_max * 1e-3), 'unit': 'samples/s', 'ms_per_step': e2e_max / steps,
                  'h2d_bytes_per_step': int(2 * mags_h[0].numel() * 4), 'd2h_bytes_per_step': 16}
  del eng
  torch.cuda.empty_cache()
  return out


def run_gpu_train(args):
  import torch
  import torch.distributed as dist

  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  with ClockSampler(local) as clk:
    r = measure_train_step(args.model, args.batch, world, rank, args.steps, args.warmup, barrier,
                           overlap=not args.no_overlap, use_graphs=not args.no_graph)
  peaks = _peaks()
  tf32_peak = _tf32_peak()
  line = {
      'metric': 'G+D train-step samples/sec', 'value': r['value'], 'unit': 'samples/s', 'n_gpus': world,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'],
      'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': 'tf32 fwd/dgrad (fp32 accumulate), fp32 wgrad/optimizer', 'data': 'synthetic',
      'config': _train_config(args, world),
      'e2e': r['e2e'],
      'gpu_launches': int(r['launches_per_step'] * args.steps * 2), 'launches_per_step': r['launches_per_step'],
      'clocks': clk.summary(),
      'roofline': {'kernel': 'whole step (all conv GEMMs)', 'bound': 'tensor',
                   'achieved': r['tflops_per_gpu'], 'peak': tf32_peak, 'unit': 'TFLOP/s',
                   'frac': r['tflops_per_gpu'] / tf32_peak, 'traffic': None,
                   'peak_source': 'cuBLAS TF32 8192^3 measured in this run (convs compute in kind::tf32); '
                                  'MEASURED_PEAKS.json bf16 sustained = %.1f' % peaks['tensor_sustained']},
      'train_step': r,
  }
  if rank == 0:
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def run_reference_train(args):
  import torch
  from oracle import nets_torch as O
  from oracle import spectral_np as OS
  import numpy as np
  if int(os.environ.get('RANK', '0')) != 0:
    return
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  spec = O.SMALL if args.model == 'small' else O.REGULAR
  P = O.init_params(spec, seed=0)
  W = torch.from_numpy(OS.create_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80).astype(np.float32))
  Winv = torch.from_numpy(OS.create_inverse_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80).astype(np.float32))
  sample = 2 if args.model == 'small' else 1
  mags = [_synthetic_mag(sample, 100 + i) for i in range(2)]
  opt_d, opt_g = O.TFAdam(O.d_names(P), P), O.TFAdam(O.g_names(P), P)
  times = []
  for it in range(max(args.warmup, 1) + max(args.steps, 1)):
    t0 = time.perf_counter()
    xs = [((m[..., 0] @ W.t()) @ Winv.t()).unsqueeze(-1) for m in mags]
    O.train_step(P, opt_d, opt_g, (xs[0], mags[0]), (xs[1], mags[1]), spec)
    if it >= max(args.warmup, 1):
      times.append(time.perf_counter() - t0)
  sps = sample / (sum(times) / len(times))
  line = {'impl': 'reference', 'metric': 'G+D train-step samples/sec', 'value': sps, 'unit': 'samples/s',
          'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
          'ms_per_step': 1e3 * args.batch / sps, 'higher_is_better': True, 'scaling': 'weak',
          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': _train_config(args, 1),
          'cpu_baseline': {'value': sps, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                           'sample': 'oracle/nets_torch.py train_step (PyTorch-CPU autograd + TF1 Adam), '
                                     '%d of %d samples per step, dropout off' % (sample, args.batch)},
          'e2e': {'value': sps, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
          'gpu_launches': 0}
  print(json.dumps(line))


# ---------------------------------------------------------------------------
# spectral workload (BASELINE configs[0] + SURVEY 8(d) config 1): waveform -> r9y9 mel
# ---------------------------------------------------------------------------
def run_gpu_spectral(args):
  """value: mel frames/s of the fused STFT->|.|->mel->dB kernel on a resident batch of
  [B, 65536] waveforms; e2e: the numpy entry point `waveform_to_r9y9_melspec` on
  the reference's own case (1 s of 22.05 kHz audio -> 87 frames, launch-bound), host buffers.
  Beside it, as a comparison leg only (north_star: "a cuFFT-free radix kernel where it beats a batched
  cuFFT call"): the same features from library calls -- frame gather + window, batched cuFFT R2C
  (torch.fft.rfft), |.|, cuBLAS matmul with the mel filterbank, dB + clip -- on the same input, and the
  fused kernel at a batch large enough for HBM to matter (512 x 65536 samples = 134 MB in)."""
  import numpy as np
  import torch
  from advoc_b200 import _native as N
  from advoc_b200 import spectral
  from oracle import spectral_np as OS
  torch.cuda.set_device(0)
  B, n = args.batch, 65536
  g = torch.Generator().manual_seed(3)
  wav = (torch.rand(B, n, 1, 1, generator=g) * 2 - 1).cuda()
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  stream = torch.cuda.current_stream()
  n0 = N.launch_count()
  out = spectral.waveform_to_r9y9_melspec_tf(wav)
  per_step = N.launch_count() - n0
  frames = out.shape[0] * out.shape[1]

  def timed(fn, steps, warmup):
    for _ in range(warmup):
      fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    torch.cuda.synchronize()
    for a, b in evs:
      flush.zero_()
      a.record(stream)
      fn()
      b.record(stream)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs)

  # ---- comparison leg: cuFFT + cuBLAS through torch (library code, not the product path) ----
  win = torch.from_numpy(spectral.lws_hann_default(1024, 256, np.float32)).cuda()
  fb = torch.from_numpy(spectral.create_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80)
                        .astype(np.float32)).cuda()
  min_level = float(np.exp(-100 / 20 * np.log(10)))

  def cufft_path(w):
    x = w[:, :, 0, 0]
    x = torch.nn.functional.pad(x, (0, 768))                      # pad_end: ceil(n / hop) frames
    fr = x.unfold(1, 1024, 256) * win                             # frame gather + window (materialised)
    mag = torch.fft.rfft(fr, dim=2).abs()                         # batched cuFFT R2C
    mel = mag @ fb.t()                                            # cuBLAS
    db = 20.0 * torch.log10(torch.clamp_min(mel, min_level)) - 20.0
    return torch.clamp((db + 100.0) / 100.0, 0.0, 1.0)

  ref_lib = cufft_path(wav)
  lib_parity = float((out[..., 0] - ref_lib).norm() / ref_lib.norm())

  with ClockSampler(0) as clk:
    total_ms = timed(lambda: spectral.waveform_to_r9y9_melspec_tf(wav), args.steps, args.warmup)
    lib_ms = timed(lambda: cufft_path(wav), args.steps, args.warmup)
    # large batch: 512 waveforms (134 MB of samples), where the input no longer fits on-chip
    Bb = 512
    wav_big = (torch.rand(Bb, n, 1, 1, generator=torch.Generator().manual_seed(4)) * 2 - 1).cuda()
    big_ms = timed(lambda: spectral.waveform_to_r9y9_melspec_tf(wav_big), max(args.steps // 5, 3), 3) / max(args.steps // 5, 3)
    big_lib_ms = timed(lambda: cufft_path(wav_big), max(args.steps // 5, 3), 3) / max(args.steps // 5, 3)
  value = frames * args.steps / (total_ms * 1e-3)
  rng = np.random.RandomState(0)
  x1 = rng.uniform(-1, 1, 22050).astype(np.float32).reshape(-1, 1, 1)
  for _ in range(3):
    spectral.waveform_to_r9y9_melspec(x1)
  t0 = time.perf_counter()
  reps = 50
  for _ in range(reps):
    y1 = spectral.waveform_to_r9y9_melspec(x1)
  e2e_s = (time.perf_counter() - t0) / reps
  t0 = time.perf_counter()
  for _ in range(5):
    ref = OS.waveform_to_r9y9_melspec(x1)
  cpu_s = (time.perf_counter() - t0) / 5
  peaks = _peaks()
  bytes_per_step = 4.0 * B * n + 4.0 * 80 * frames
  ach = bytes_per_step * args.steps / (total_ms * 1e-3) / 1e9
  frames_big = Bb * 256
  bytes_big = 4.0 * Bb * n + 4.0 * 80 * frames_big
  ach_big = bytes_big / (big_ms * 1e-3) / 1e9
  # FFT arithmetic (5 N log2 N per complex FFT of N points; two real frames share one complex FFT) + window
  flop_per_frame = 0.5 * 5 * 1024 * 10 + 1024 + 2 * 513 * 2
  true_bound = _ncu_issue_bound('stft_pair_kernel')
  line = {
      'metric': 'mel-frames/sec STFT+mel', 'value': value, 'unit': 'mel-frames/s', 'n_gpus': 1, 'steps': args.steps,
      'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
      'config': {'workload': 'advoc.spectral r9y9 mel (nfft 1024, hop 256, 80 mels) on uniform noise [%d, 65536]; '
                             'e2e = the reference case 1x22050 samples -> (87, 80, 1) through the numpy API' % B,
                 'l2': 'flushed between timed steps'},
      'e2e': {'value': 87 / e2e_s, 'unit': 'mel-frames/s', 'us_per_call': e2e_s * 1e6,
              'h2d_bytes_per_step': 88200, 'd2h_bytes_per_step': 87 * 80 * 8,
              'note': 'one 87-frame call: launch / copy latency bound (2 kernels, 2 small copies), not a throughput figure'},
      'gpu_launches': int(per_step * args.steps), 'launches_per_step': int(per_step), 'clocks': clk.summary(),
      'roofline': {'kernel': 'stft_pair_kernel', 'bound': 'hbm', 'achieved': ach_big, 'peak': peaks['hbm'], 'unit': 'GB/s',
                   'frac': ach_big / peaks['hbm'], 'traffic': None, 'peak_source': peaks['source'] + ' copy bandwidth',
                   'bytes_per_launch': bytes_big, 'ms_per_launch': big_ms,
                   'at': '[512 x 65536] samples (the HBM roofline is quoted at the batch where the input exceeds L2)',
                   'at_batch_%d' % B: {'achieved': ach, 'frac': ach / peaks['hbm'], 'bytes_per_launch': bytes_per_step},
                   'fp32_tflops': flop_per_frame * frames_big / (big_ms * 1e-3) / 1e12,
                   'true_bound': true_bound,
                   'note': 'the kernel is NOT HBM-bound at either batch (DRAM 1.8-2.6 % busy): every sample is re-read '
                           'four times from L1/L2 (hop = nfft / 4) and each frame costs ~27 kFLOP of shared-memory radix-4 '
                           'butterflies; ncu shows it bound by instruction issue (true_bound: sm__issue_active of the '
                           'committed capture), i.e. by the CUDA-core FFT itself'},
      'vs_cufft': {'fused_kernel_ms': total_ms / args.steps, 'cufft_cublas_ms': lib_ms / args.steps,
                   'speedup': lib_ms / total_ms,
                   'at_512x65536': {'fused_kernel_ms': big_ms, 'cufft_cublas_ms': big_lib_ms, 'speedup': big_lib_ms / big_ms},
                   'parity_rel_l2_vs_library_path': lib_parity,
                   'library_path': 'pad + unfold * window (materialised frames), torch.fft.rfft (batched cuFFT R2C), abs, '
                                   'cuBLAS matmul with the mel filterbank, log10 / clip -- comparison only, never on the product path'},
      'cpu_baseline': {'value': 87 / cpu_s, 'unit': 'mel-frames/s', 'cores': 1, 'kind': 'port',
                       'sample': 'oracle/spectral_np.py on the 22050-sample case, mean of 5 (us/call %.0f)' % (cpu_s * 1e6)},
      'parity_rel_l2': float(np.linalg.norm(y1 - ref) / np.linalg.norm(ref)),
  }
  print(json.dumps(line))


# ---------------------------------------------------------------------------
# MelspecGAN workload (BASELINE configs[4]): one outer iteration of models/melspecgan/train.py:149-153
# ---------------------------------------------------------------------------
MELSPECGAN_GFLOP = {'G_f': 0.3994, 'D_f': 0.3973}


def run_gpu_melspecgan(args):
  import torch
  import torch.distributed as dist
  from advoc_b200 import _native as N
  from advoc_b200 import melspecgan as MG
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
  B = args.batch
  eng = MG.MelspecGAN(MG.init_params(seed=0), B, train_loss=args.melspecgan_loss, world_size=world)
  n_d = 1 if args.melspecgan_loss == 'dcgan' else 5
  g = torch.Generator().manual_seed(10 + rank)
  xs_h = [(torch.rand(B, 64, 80, 1, generator=g) * 2 - 1).pin_memory() for _ in range(n_d)]
  xs_d = [x.cuda() for x in xs_h]
  stage = [torch.empty_like(x) for x in xs_d]
  zs = [torch.randn(B, MG.Z_DIM, generator=g).cuda() for _ in range(n_d + 1)]
  stream = torch.cuda.current_stream()

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def step_device():
    eng.train_loop(xs_d, zs[:n_d], zs[n_d])

  def step_e2e():
    for s, h in zip(stage, xs_h):
      s.copy_(h, non_blocking=True)
    eng.train_loop(stage, zs[:n_d], zs[n_d])
    return eng.loss_values()

  def timed(fn, steps, warmup):
    for _ in range(warmup):
      fn()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
      fn()
    b.record(stream)
    barrier()
    return a.elapsed_time(b)

  # kernels of one steady-state outer iteration: launches the library counted directly plus the kernels
  # inside the CUDA graphs the engine replayed (recorded launch counts per graph), minus launches that
  # were only recorded during a capture
  step_device()
  torch.cuda.synchronize()
  n0, r0, c0 = N.launch_count(), eng.replayed_launches, eng.capture_launches
  step_device()
  torch.cuda.synchronize()
  per_step = (N.launch_count() - n0) - (eng.capture_launches - c0) + (eng.replayed_launches - r0)
  with ClockSampler(local) as clk:
    total_ms = timed(step_device, args.steps, args.warmup)
    e2e_ms = timed(step_e2e, args.steps, args.warmup)
  if world > 1:
    t = torch.tensor([total_ms, e2e_ms], device='cuda', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])
  samples = world * B * args.steps
  value = samples / (total_ms * 1e-3)
  # D step = G_f + 2 D_f + 2 (2 D_f) [+ 6 D_f for the penalty: forward, first backward (dgrad), the
  # adjoint sweep (forward convs + wgrads) and the seeded backward (dgrad + wgrad)]; G step = 3 G_f + 2 D_f
  gf, df = MELSPECGAN_GFLOP['G_f'], MELSPECGAN_GFLOP['D_f']
  gflop = n_d * (gf + (12 if args.melspecgan_loss == 'wgangp' else 6) * df) + (3 * gf + 2 * df)
  tf32_peak = _tf32_peak()
  tflops = gflop * 1e9 * value / 1e12
  line = {
      'metric': 'MelspecGAN outer-iteration samples/sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world,
      'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True,
      'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32 convs (fp32 accumulate), fp32 elsewhere',
      'data': 'synthetic',
      'config': {'workload': 'MelspecGAN 64x80 train iteration, train_loss=%s (%d D steps + 1 G step, '
                             'models/melspecgan/train.py:149-153), batch %d per GPU' % (args.melspecgan_loss, n_d, B),
                 'batch_per_gpu': B, 'global_batch': B * world, 'parallelism': 'dp%d' % world,
                 'l2': 'working set per iteration exceeds L2; no explicit flush',
                 'launch': 'd_step / g_step replayed from captured CUDA graphs' if eng.use_graphs else 'eager'},
      'e2e': {'value': samples / (e2e_ms * 1e-3), 'unit': 'samples/s', 'ms_per_step': e2e_ms / args.steps,
              'h2d_bytes_per_step': int(n_d * xs_h[0].numel() * 4), 'd2h_bytes_per_step': 8},
      'gpu_launches': int(per_step * args.steps * 2), 'launches_per_step': int(per_step), 'clocks': clk.summary(),
      'roofline': {'kernel': 'whole iteration (conv GEMMs)', 'bound': 'tensor', 'achieved': tflops / world,
                   'peak': tf32_peak, 'unit': 'TFLOP/s', 'frac': tflops / world / tf32_peak, 'traffic': None,
                   'peak_source': 'cuBLAS TF32 8192^3 measured in this run'},
  }
  if rank == 0:
    print(json.dumps(line))
  if world > 1:
    dist.destroy_process_group()


def run_reference_melspecgan(args):
  """CPU restatement of one MelspecGAN outer iteration (oracle/melspecgan_torch.py, autograd incl. the
  double backward of the gradient penalty) on a bounded sample of the batch."""
  import torch
  from oracle import melspecgan_torch as M
  from oracle.nets_torch import TFAdam
  if int(os.environ.get('RANK', '0')) != 0:
    return
  cores = os.cpu_count() or 1
  torch.set_num_threads(cores)
  loss = args.melspecgan_loss
  n_d = 1 if loss == 'dcgan' else 5
  sample = 8
  P = M.init_params(seed=0)
  g = torch.Generator().manual_seed(10)
  dn, gn = M.d_names(P), M.g_names(P)
  lr, b2 = (2e-4, 0.999) if loss == 'dcgan' else (1e-4, 0.9)
  opt_d, opt_g = TFAdam(dn, P, lr=lr, beta2=b2), TFAdam(gn, P, lr=lr, beta2=b2)

  def grads(names, which, z, x, alpha):
    Pr = {n: (t.clone().requires_grad_(True) if n in names else t) for n, t in P.items()}
    if loss == 'wgan':
      G_z = M.generator(Pr, z)
      D_x, D_G_z = M.discriminator(Pr, x), M.discriminator(Pr, G_z)
      l = dict(D_loss=D_G_z.mean() - D_x.mean(), G_loss=-D_G_z.mean())
    else:
      l = M.losses(Pr, z, x, loss, alpha=alpha)
    gs = torch.autograd.grad(l[which], [Pr[n] for n in names], allow_unused=True)
    return {n: (gv if gv is not None else torch.zeros_like(P[n])) for n, gv in zip(names, gs)}

  times = []
  for it in range(max(args.warmup, 1) + max(args.steps, 1)):
    t0 = time.perf_counter()
    for _ in range(n_d):
      x = torch.rand(sample, 64, 80, 1, generator=g) * 2 - 1
      z = torch.randn(sample, M.Z_DIM, generator=g)
      opt_d.step(P, grads(dn, 'D_loss', z, x, torch.rand(sample, 1, 1, 1, generator=g)))
    z = torch.randn(sample, M.Z_DIM, generator=g)
    opt_g.step(P, grads(gn, 'G_loss', z, torch.rand(sample, 64, 80, 1, generator=g) * 2 - 1,
                        torch.rand(sample, 1, 1, 1, generator=g)))
    if it >= max(args.warmup, 1):
      times.append(time.perf_counter() - t0)
  sps = sample / (sum(times) / len(times))
  line = {'impl': 'reference', 'metric': 'MelspecGAN outer-iteration samples/sec', 'value': sps, 'unit': 'samples/s',
          'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * args.batch / sps,
          'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': {'workload': 'MelspecGAN 64x80 train iteration, train_loss=%s (%d D steps + 1 G step), batch %d'
                                 % (loss, n_d, args.batch)},
          'cpu_baseline': {'value': sps, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
                           'sample': 'oracle/melspecgan_torch.py with torch autograd + TF1 Adam, %d of %d samples '
                                     'per step' % (sample, args.batch)},
          'e2e': {'value': sps, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
          'gpu_launches': 0}
  print(json.dumps(line))


def run_reference_spectral(args):
  import numpy as np
  from oracle import spectral_np as OS
  if int(os.environ.get('RANK', '0')) != 0:
    return
  rng = np.random.RandomState(0)
  x = rng.uniform(-1, 1, (4, 65536)).astype(np.float32)
  times = []
  for it in range(max(args.warmup, 1) + max(args.steps, 1)):
    t0 = time.perf_counter()
    frames = 0
    for i in range(x.shape[0]):
      frames += OS.waveform_to_r9y9_melspec(x[i].reshape(-1, 1, 1)).shape[0]
    if it >= max(args.warmup, 1):
      times.append(time.perf_counter() - t0)
  fps = frames / (sum(times) / len(times))
  line = {'impl': 'reference', 'metric': 'mel-frames/sec STFT+mel', 'value': fps, 'unit': 'mel-frames/s',
          'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
          'ms_per_step': 1e3 * args.batch * 256 / fps, 'higher_is_better': True, 'scaling': 'weak',
          'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
          'config': {'workload': 'advoc.spectral r9y9 mel on uniform noise [%d, 65536]' % args.batch},
          'cpu_baseline': {'value': fps, 'unit': 'mel-frames/s', 'cores': 1, 'kind': 'port',
                           'sample': 'oracle/spectral_np.py (numpy f64, single thread like lws), 4 of %d waveforms '
                                     'per step' % args.batch},
          'e2e': {'value': fps, 'unit': 'mel-frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
          'gpu_launches': 0}
  print(json.dumps(line))


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=50)
  ap.add_argument('--warmup', type=int, default=5)
  ap.add_argument('--impl', default='advoc_b200', choices=['advoc_b200', 'reference'])
  ap.add_argument('--model', default='small', choices=['small', 'regular'])
  ap.add_argument('--batch', type=int, default=32, help='samples per GPU')
  ap.add_argument('--workload', default='infer', choices=['infer', 'train', 'spectral', 'melspecgan'])
  ap.add_argument('--melspecgan-loss', default='wgangp', choices=['dcgan', 'wgan', 'wgangp'])
  ap.add_argument('--math', default='f16', choices=['f16', 'tf32', 'fp32'],
                  help='generator-forward operand type: f16 (default) or tf32 (fp32 storage)')
  ap.add_argument('--no-graph', action='store_true')
  ap.add_argument('--no-overlap', action='store_true', help='train step: blocking all-reduces (A/B)')
  ap.add_argument('--no-train-step', action='store_true',
                  help='default line: skip the data-parallel train-step measurement')
  ap.add_argument('--no-cpu-baseline', action='store_true')
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl != 'reference' else args.warmup
  if args.workload == 'spectral':
    run_reference_spectral(args) if args.impl == 'reference' else run_gpu_spectral(args)
  elif args.workload == 'melspecgan':
    if args.batch == 32:
      args.batch = 64          # models/melspecgan/train.py:13
    run_reference_melspecgan(args) if args.impl == 'reference' else run_gpu_melspecgan(args)
  elif args.workload == 'train':
    if args.impl == 'reference':
      run_reference_train(args)
    else:
      run_gpu_train(args)
  elif args.impl == 'reference':
    run_reference(args)
  else:
    run_gpu(args)


if __name__ == '__main__':
  main()
