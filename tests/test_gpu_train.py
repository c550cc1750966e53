"""GPU parity of the G+D train step (forward, backward, TF1-Adam) against autograd on the
PyTorch-CPU oracle (oracle/nets_torch.py; parity unpinned by the reference).  The exact-fp32
math mode checks the backward-pass construction tightly (2e-4); the TF32 tensor-core mode is
held to 1e-2 on gradients (two passes through ~20 TF32 layers; bias gradients are sums
with heavy cancellation and land near 5e-3) and 1e-3 on losses."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
  a = a.detach().double().cpu()
  b = b.detach().double().cpu()
  return float((a - b).norm() / b.norm().clamp_min(1e-300))


def _setup(math, seed=0, batch=1):
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from advoc_b200.train import TrainEngine
  from oracle import nets_torch as O
  P = O.init_params(O.SMALL, seed=seed)
  g = torch.Generator().manual_seed(11)
  for k in P:
    if k.endswith('/bias'):
      P[k] = torch.randn(P[k].shape, generator=g) * 0.05
  spec = nets.GenSpec(32, 5, (5, 4))
  eng = TrainEngine(spec, 32, {k: v.cuda() for k, v in P.items()}, batch,
                    math=N.MATH_FP32 if math == 'fp32' else N.MATH_AUTO)
  target = torch.randn(batch, 256, 513, 1, generator=g).abs() * 0.1
  x = (target + torch.randn(batch, 256, 513, 1, generator=g) * 0.02)
  masks = {k: (torch.rand(eng.G.dropout_shape(k), generator=g) < 0.5) for k in (5, 4)}
  full = {k: torch.cat([m, torch.zeros(m.shape[0], m.shape[1], 1, m.shape[3], dtype=torch.bool)], 2).float()
          for k, m in masks.items()}
  dmasks = {k: m.to(torch.uint8).cuda().contiguous() for k, m in masks.items()}
  return O, P, eng, x, target, full, dmasks


@pytest.mark.parametrize('math,gtol', [('fp32', 2e-4), ('auto', 1e-2)])
def test_d_step_gradients(math, gtol):
  O, P, eng, x, target, full, dmasks = _setup(math)
  eng.d_step(x.cuda(), target.cuda(), dropout=dmasks, apply=False)
  Pd = {n: t.clone().requires_grad_(n.startswith('discriminator')) for n, t in P.items()}
  l = O.losses(Pd, x, target, O.SMALL, dropout_masks=full)
  ref = O.grads_of(l['d_loss'], Pd, O.d_names(P))
  assert abs(eng.loss_values()[0] - float(l['d_loss'])) < 1e-3 * abs(float(l['d_loss']))
  for n in O.d_names(P):
    assert _rel(eng.flat.G[n], ref[n]) < gtol, n


@pytest.mark.parametrize('math,gtol', [('fp32', 2e-4), ('auto', 1e-2)])
def test_g_step_gradients(math, gtol):
  O, P, eng, x, target, full, dmasks = _setup(math)
  eng.g_step(x.cuda(), target.cuda(), dropout=dmasks, apply=False)
  Pg = {n: t.clone().requires_grad_(n.startswith('generator')) for n, t in P.items()}
  l = O.losses(Pg, x, target, O.SMALL, dropout_masks=full)
  ref = O.grads_of(l['g_loss'], Pg, O.g_names(P))
  _, g_gan, g_l1 = eng.loss_values()
  assert abs(g_gan - float(l['g_gan'])) < 1e-3 * abs(float(l['g_gan']))
  assert abs(g_l1 - 10.0 * float(l['g_l1'])) < 1e-3 * abs(10.0 * float(l['g_l1']))
  for n in O.g_names(P):
    assert _rel(eng.flat.G[n], ref[n]) < gtol, n


def test_train_loop_matches_oracle_adam():
  """Two reference `train_loop`s (D step then G step, separate minibatches) in exact-fp32 mode:
  parameters after TF1-Adam follow the oracle trajectory."""
  O, P, eng, x, target, full, dmasks = _setup('fp32')
  g = torch.Generator().manual_seed(5)
  x2 = x + torch.randn(x.shape, generator=g) * 0.01
  t2 = target * 1.1
  opt_d = O.TFAdam(O.d_names(P), P)
  opt_g = O.TFAdam(O.g_names(P), P)
  Pref = dict(P)
  for it in range(2):
    step = eng.train_loop((x.cuda(), target.cuda()), (x2.cuda(), t2.cuda()), dropout=dmasks)
    O.train_step(Pref, opt_d, opt_g, (x, target), (x2, t2), O.SMALL, full, full)
    assert step == it + 1
  for n in P:
    # Adam's first steps move every weight by ~lr regardless of the gradient scale, so compare
    # the UPDATE (p - p0), not p
    upd, ref = eng.P[n].cpu() - P[n], Pref[n] - P[n]
    assert _rel(upd, ref) < 2e-2, n


def test_resume_from_adam_slots():
  """Engine restarted from parameters + Adam slots (the `/Adam`, `/Adam_1` variables of a TF training
  checkpoint, advoc_b200.checkpoint.load_adam_slots) continues like an uninterrupted run (to the
  summation-order noise of the atomically accumulated filter gradients)."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from advoc_b200.train import TrainEngine
  O, P, eng, x, target, full, dmasks = _setup('fp32')
  batch = (x.cuda(), target.cuda())
  eng.train_loop(batch, batch, dropout=dmasks)
  snap_p = {n: t.clone() for n, t in eng.P.items()}
  fl = eng.flat
  snap_m = {n: fl.m[fl.offsets[n]:fl.offsets[n] + fl.P[n].numel()].view(fl.P[n].shape).clone() for n in fl.names}
  snap_v = {n: fl.v[fl.offsets[n]:fl.offsets[n] + fl.P[n].numel()].view(fl.P[n].shape).clone() for n in fl.names}
  eng.train_loop(batch, batch, dropout=dmasks)
  eng2 = TrainEngine(nets.GenSpec(32, 5, (5, 4)), 32, snap_p, 1, math=N.MATH_FP32)
  eng2.restore_adam(snap_m, snap_v, 1)
  assert eng2.train_loop(batch, batch, dropout=dmasks) == 2
  for n in P:
    assert float((eng2.P[n] - eng.P[n]).abs().max()) < 2e-6, n   # one Adam step moves a weight by ~2e-4
  with pytest.raises(KeyError):
    eng2.restore_adam({}, snap_v, 1)


@pytest.mark.parametrize('cin,cout,k', [(1, 32, 4), (1, 64, 4), (1, 128, 4), (2, 32, 4), (2, 64, 4), (1, 64, 5),
                                        (1, 16, 4)])
def test_thin_filter_gradients(cin, cout, k):
  """advoc_conv2d_wgrad on the one/two-channel-input layers (encoder_1, discriminator layer_1,
  decoder_1, MelspecGAN conv_0 / upconv_4): shared-memory tiled kernels and their fallbacks against
  torch's conv2d weight gradient (fp32, odd width, stride 2, SAME-style padding 1)."""
  import ctypes as C
  import torch.nn.functional as F
  from advoc_b200 import _native as N
  from advoc_b200.nets import _ptr, _stream
  g = torch.Generator().manual_seed(cin * 100 + cout + k)
  B, H, W = 3, 20, 37
  ho, wo = (H + 1 + (2 if k == 5 else 1) - k) // 2 + 1, (W + 1 + 2 - k) // 2 + 1
  x = torch.randn(B, H, W, cin, generator=g)
  dy = torch.randn(B, ho, wo, cout, generator=g)
  w = torch.zeros(k, k, cin, cout, requires_grad=True)
  xt = F.pad(x.permute(0, 3, 1, 2), (1, 2, 1, 2))
  y = F.conv2d(xt, w.permute(3, 2, 0, 1), stride=2)[:, :, :ho, :wo]
  (ref,) = torch.autograd.grad(y, w, dy.permute(0, 3, 1, 2))
  d = N.ConvDesc(B, H, W, cin, cout, k, k, 2, 2, 1, 1, ho, wo, N.MATH_FP32)
  dw = torch.zeros(k, k, cin, cout, device='cuda')
  xd, dyd = x.cuda().contiguous(), dy.cuda().contiguous()
  N.call('advoc_conv2d_wgrad', C.byref(d), _ptr(xd), cin, _ptr(dyd), cout, _ptr(dw), _stream())
  assert _rel(dw, ref) < 1e-4


# ---------------------------------------------------------------------------------------------
# round 2: the benchmarked configurations (regular model, TF32 trajectory) and real NCCL
# ---------------------------------------------------------------------------------------------
def _setup_regular(math, batch=1):
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from advoc_b200.train import TrainEngine
  from oracle import nets_torch as O
  P = O.init_params(O.REGULAR, seed=0)
  g = torch.Generator().manual_seed(12)
  for k in P:
    if k.endswith('/bias'):
      P[k] = torch.randn(P[k].shape, generator=g) * 0.05
  spec = nets.GenSpec(64, 8, (8, 7, 6))
  eng = TrainEngine(spec, 64, {k: v.cuda() for k, v in P.items()}, batch,
                    math=N.MATH_FP32 if math == 'fp32' else N.MATH_AUTO)
  target = torch.randn(batch, 256, 513, 1, generator=g).abs() * 0.1
  x = (target + torch.randn(batch, 256, 513, 1, generator=g) * 0.02)
  return O, P, eng, x, target


@pytest.mark.parametrize('math,tol_all,tol_each', [('fp32', 1e-3, 5e-3), ('auto', 1e-2, 6e-2)])
def test_regular_model_step_gradients(math, tol_all, tol_each):
  """BASELINE configs[2] net (AdVoc regular: ngf = ndf = 64, 8 + 8 layers, 1x3 bottleneck, 1024-channel
  concats): D-step and G-step gradients on the exact-fp32 path (checks the backward construction) and on
  the production TF32 path, against FLOAT64 autograd on the oracle.

  Why float64 and why two tolerances: at initialisation every patch probability is ~0.5, so the GAN
  gradient entering the generator is a near-constant field plus a small informative part, and that part is
  ill-conditioned in single precision -- torch's own CPU fp32 autograd is 5e-4 off float64 on d loss_GAN /
  d generated (scripts/dev_dis_bwd_diag.py, profiles/README.md).  The error is white noise injected at
  full resolution: it averages out in the filter gradients of the shallow, many-pixel layers (encoder_1:
  3e-6) and grows ~2x per level towards the 1x3 bottleneck (encoder_8: 2e-3 on the fp32 path), whose
  gradients are 1000x smaller in norm.  `tol_all` bounds the relative L2 error of the WHOLE generator
  gradient (what Adam sees), `tol_each` every single tensor."""
  O, P, eng, x, target = _setup_regular(math)
  P64 = {n: t.double() for n, t in P.items()}
  x64, t64 = x.double(), target.double()
  eng.d_step(x.cuda(), target.cuda(), dropout=None, apply=False)
  Pd = {n: t.clone().requires_grad_(n.startswith('discriminator')) for n, t in P64.items()}
  l = O.losses(Pd, x64, t64, O.REGULAR)
  ref = O.grads_of(l['d_loss'], Pd, O.d_names(P))
  assert abs(eng.loss_values()[0] - float(l['d_loss'])) < 1e-3 * abs(float(l['d_loss']))
  for n in O.d_names(P):
    assert _rel(eng.flat.G[n], ref[n]) < (5e-4 if math == 'fp32' else 1e-2), n
  eng.g_step(x.cuda(), target.cuda(), dropout=None, apply=False)
  Pg = {n: t.clone().requires_grad_(n.startswith('generator')) for n, t in P64.items()}
  l = O.losses(Pg, x64, t64, O.REGULAR)
  ref = O.grads_of(l['g_loss'], Pg, O.g_names(P))
  _, g_gan, g_l1 = eng.loss_values()
  assert abs(g_gan - float(l['g_gan'])) < 1e-3 * abs(float(l['g_gan']))
  assert abs(g_l1 - 10.0 * float(l['g_l1'])) < 1e-3 * abs(10.0 * float(l['g_l1']))
  errs = {n.replace('generator/', '').replace('/conv2d_transpose', '').replace('/conv2d', ''):
              round(_rel(eng.flat.G[n], ref[n]), 5) for n in O.g_names(P)}
  got_all = torch.cat([eng.flat.G[n].reshape(-1).double().cpu() for n in O.g_names(P)])
  ref_all = torch.cat([ref[n].reshape(-1) for n in O.g_names(P)])
  whole = float((got_all - ref_all).norm() / ref_all.norm())
  assert whole < tol_all, (whole, errs)
  assert max(errs.values()) < tol_each, errs


def test_train_loop_trajectory_tf32():
  """Two reference `train_loop`s on the production path (TF32 tensor cores, the mode bench.py times):
  the parameter UPDATES follow the oracle's TF1-Adam trajectory."""
  O, P, eng, x, target, full, dmasks = _setup('auto')
  g = torch.Generator().manual_seed(5)
  x2 = x + torch.randn(x.shape, generator=g) * 0.01
  t2 = target * 1.1
  opt_d = O.TFAdam(O.d_names(P), P)
  opt_g = O.TFAdam(O.g_names(P), P)
  Pref = dict(P)
  for it in range(2):
    step = eng.train_loop((x.cuda(), target.cuda()), (x2.cuda(), t2.cuda()), dropout=dmasks)
    O.train_step(Pref, opt_d, opt_g, (x, target), (x2, t2), O.SMALL, full, full)
    assert step == it + 1
  for n in P:
    upd, ref = eng.P[n].cpu() - P[n], Pref[n] - P[n]
    assert _rel(upd, ref) < 5e-2, (n, _rel(upd, ref))


@pytest.mark.parametrize('overlap', ['1', '0'])
def test_two_rank_nccl_train_loop_matches_accumulated_gradients(tmp_path, overlap):
  """Real NCCL: two ranks run one data-parallel `train_loop` (one sample each, D all-reduce deferred
  under the generator forward, G all-reduce in three buckets) and must land on the parameters a single
  engine reaches by accumulating the two samples' gradients and applying Adam on their mean
  (SURVEY 8(e) equivalence).  Skipped on a one-GPU box."""
  import os
  import subprocess
  import sys
  if torch.cuda.device_count() < 2:
    pytest.skip('needs 2 GPUs')
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from advoc_b200.train import TrainEngine
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  sys.path.insert(0, os.path.join(root, 'tests'))
  import dp_worker
  env = dict(os.environ)
  env.update(DP_MATH='fp32', DP_OVERLAP=overlap)
  r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                      '--master-addr', '127.0.0.1', '--master-port', '29731', os.path.join(root, 'tests', 'dp_worker.py'),
                      str(tmp_path)], cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                     universal_newlines=True)
  assert r.returncode == 0, r.stdout[-3000:]
  got = [torch.load(os.path.join(str(tmp_path), 'rank%d.pt' % k)) for k in range(2)]
  assert torch.equal(got[0]['p'], got[1]['p'])          # replicas stay in lock step
  # single engine: accumulate the two samples' gradients, Adam on the mean, D first then G
  eng = TrainEngine(nets.GenSpec(32, 5, (5, 4)), 32, nets.init_params(32, 32, 5, seed=0), 1, math=N.MATH_FP32)
  f = eng.flat
  data = [dp_worker.data(k) for k in range(2)]
  lo, hi = f.dis_range()
  acc = torch.zeros_like(f.g)
  for x, tgt in data:
    eng.d_step(x[0].cuda(), tgt[0].cuda(), dropout=None, apply=False)
    acc[lo:hi] += f.g[lo:hi]
  f.g[lo:hi] = acc[lo:hi] * 0.5
  eng._adam(lo, hi, 1)
  eng.refresh_weights('D')
  lo, hi = f.gen_range()
  for x, tgt in data:
    eng.g_step(x[1].cuda(), tgt[1].cuda(), dropout=None, apply=False)
    acc[lo:hi] += f.g[lo:hi]
  f.g[lo:hi] = acc[lo:hi] * 0.5
  eng._adam(lo, hi, 1)
  p0 = TrainEngine(nets.GenSpec(32, 5, (5, 4)), 32, nets.init_params(32, 32, 5, seed=0), 1, math=N.MATH_FP32).flat.p.cpu()
  upd_got, upd_want = got[0]['p'] - p0, f.p.cpu() - p0
  assert _rel(got[0]['g'][lo:hi], acc[lo:hi].cpu()) < 1e-4      # all-reduced G gradients = the accumulated sum
  assert _rel(upd_got, upd_want) < 1e-2


@pytest.mark.parametrize('cin,cout,k', [(1, 32, 4), (1, 64, 4), (1, 128, 4), (1, 256, 4), (2, 32, 4), (2, 64, 4),
                                        (2, 128, 4), (1, 64, 5), (1, 128, 5)])
def test_thin_filter_gradients_on_tensor_cores(cin, cout, k):
  """wgrad_thin_tc_kernel (MODE 0 / 1 / 3): filter gradient of a k4 conv from one / two input channels and of a
  5x5 conv from one channel (MelspecGAN conv_0) against torch's conv2d weight gradient.  The thin side enters
  as hi + lo tf32 parts; the wide side is read by TMA and truncated to tf32 by the tensor core, so with a
  tf32-representable dy the result is fp32-accurate.  Also checks += semantics and run-to-run determinism."""
  import ctypes as C
  import torch.nn.functional as F
  from advoc_b200 import _native as N
  from advoc_b200.nets import _ptr, _stream
  g = torch.Generator().manual_seed(cin * 1000 + cout + k)
  B, H, W = 3, 52, 75
  ho, wo = (H + 1 + (2 if k == 5 else 1) - k) // 2 + 1, (W + 1 + 2 - k) // 2 + 1
  x = torch.randn(B, H, W, cin, generator=g)
  dy = torch.randn(B, ho, wo, cout + 8, generator=g)
  dy = (dy.view(torch.int32) & ~0x1FFF).view(torch.float32)        # tf32-representable
  w = torch.zeros(cout, cin, k, k, dtype=torch.float64, requires_grad=True)
  xt = F.pad(x.permute(0, 3, 1, 2), (1, 2, 1, 2))
  y = F.conv2d(xt.double(), w, stride=2)[:, :, :ho, :wo]
  (ref,) = torch.autograd.grad(y, w, dy[..., 8:].double().permute(0, 3, 1, 2).contiguous())
  ref = ref.permute(2, 3, 1, 0)              # OIHW -> HWIO
  d = N.ConvDesc(B, H, W, cin, cout, k, k, 2, 2, 1, 1, ho, wo, N.MATH_AUTO)
  xd, dyd = x.cuda().contiguous(), dy.cuda().contiguous()
  base = torch.randn(k, k, cin, cout, generator=g)
  outs = []
  for _ in range(2):
    dw = base.cuda().clone()
    n0 = N.launch_count()
    N.call('advoc_conv2d_wgrad', C.byref(d), _ptr(xd), cin, _ptr(dyd[..., 8:]), cout + 8, _ptr(dw), _stream())
    assert N.launch_count() - n0 == 2          # tensor-core kernel + ordered reduction of the partials
    torch.cuda.synchronize()
    assert N.debug_flags() == 0
    outs.append(dw.cpu())
  assert torch.equal(outs[0], outs[1])
  assert _rel(outs[0] - base, ref.float()) < 1e-5


@pytest.mark.parametrize('cin', [64, 256, 512])
def test_to_one_channel_filter_gradient_on_tensor_cores(cin):
  """wgrad_thin_tc_kernel MODE 2: filter gradient of the PatchGAN head (k4 s1 pad 1 VALID, Cin -> 1)."""
  import ctypes as C
  import torch.nn.functional as F
  from advoc_b200 import _native as N
  from advoc_b200.nets import _ptr, _stream
  g = torch.Generator().manual_seed(cin)
  B, H, W = 3, 19, 37
  ho, wo = H + 2 - 4 + 1, W + 2 - 4 + 1
  x = torch.randn(B, H, W, cin, generator=g)
  x = (x.view(torch.int32) & ~0x1FFF).view(torch.float32)           # tf32-representable wide side
  dy = torch.randn(B, ho, wo, 1, generator=g)
  w = torch.zeros(1, cin, 4, 4, dtype=torch.float64, requires_grad=True)
  y = F.conv2d(x.double().permute(0, 3, 1, 2).contiguous(), w, stride=1, padding=1)
  (ref,) = torch.autograd.grad(y, w, dy.double().permute(0, 3, 1, 2).contiguous())
  ref = ref.permute(2, 3, 1, 0)              # OIHW -> HWIO
  d = N.ConvDesc(B, H, W, cin, 1, 4, 4, 1, 1, 1, 1, ho, wo, N.MATH_AUTO)
  xd, dyd = x.cuda().contiguous(), dy.cuda().contiguous()
  dw = torch.zeros(4, 4, cin, 1, device='cuda')
  n0 = N.launch_count()
  N.call('advoc_conv2d_wgrad', C.byref(d), _ptr(xd), cin, _ptr(dyd), 1, _ptr(dw), _stream())
  assert N.launch_count() - n0 == 2
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert _rel(dw, ref.float()) < 1e-5
