"""GPU parity of the MelspecGAN stacks and train step (through the C-ABI) against the PyTorch-CPU
oracle (oracle/melspecgan_torch.py; parity unpinned by the reference).  Tolerances: forward 1e-3
relative L2 on the TF32 tensor-core path (BASELINE.json north_star), 1e-4 on the exact-fp32 path
(batch-norm divides by small per-channel deviations, which amplifies fp32 summation-order noise);
gradients 6e-3 (fp32: the deepest ones pass through eight batch-normalised layers at batch 6, where ReLU gates of near-zero activations flip with summation order) / 6e-2 (TF32 operands through sixteen layer passes)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
  a = a.detach().double().cpu()
  b = b.detach().double().cpu()
  return float((a - b).norm() / b.norm().clamp_min(1e-300))


# a bias added right before a batch normalisation has an exactly zero gradient (the mean
# subtraction removes it): both sides hold rounding noise there, so compare absolutely
_DEAD_BIAS = ('D/conv_1/b', 'D/conv_2/b', 'D/conv_3/b', 'G/upconv_1/b', 'G/upconv_2/b',
              'G/upconv_3/b')


def _check_grads(eng, ref, names, gtol):
  for n in names:
    if n in _DEAD_BIAS:
      scale = float(ref[n[:-2] + '/W'].norm())
      assert float(eng.flat.G[n].norm()) < 1e-3 * scale and float(ref[n].norm()) < 1e-3 * scale, n
    else:
      # bias / beta gradients are plain sums over the batch with heavy cancellation: TF32 operand
      # rounding upstream shows up ~3x larger in them than in the filter gradients
      tol = 3 * gtol if (gtol > 1e-2 and (n.endswith('/b') or n.endswith('/beta'))) else gtol
      assert _rel(eng.flat.G[n], ref[n]) < tol, n


def _setup(math, dim, batch, loss='dcgan', seed=0):
  from advoc_b200 import _native as N
  from advoc_b200.melspecgan import MelspecGAN
  from oracle import melspecgan_torch as M
  P = M.init_params(seed=seed, dim=dim)
  g = torch.Generator().manual_seed(5)
  for n in P:   # make biases / BN parameters non-trivial
    if n.endswith('/b') or n.endswith('/beta'):
      P[n] = torch.randn(P[n].shape, generator=g) * 0.05
    if n.endswith('/gamma'):
      P[n] = 1 + torch.randn(P[n].shape, generator=g) * 0.1
  eng = MelspecGAN({k: v.cuda() for k, v in P.items()}, batch, dim=dim, train_loss=loss,
                   math=N.MATH_FP32 if math == 'fp32' else N.MATH_AUTO)
  z = torch.randn(batch, M.Z_DIM, generator=g)
  x = torch.rand(batch, 64, 80, 1, generator=g) * 2 - 1
  return M, P, eng, z, x


@pytest.mark.parametrize('math,tol', [('fp32', 1e-4), ('auto', 1e-3)])
def test_forward_matches_oracle(math, tol):
  M, P, eng, z, x = _setup(math, 64, 8)
  G_z = eng.generate(z.cuda())
  ref, layers = M.generator(P, z, 64, return_layers=True)
  assert G_z.shape == ref.shape == (8, 64, 80, 1)
  for i in range(4):
    assert _rel(eng.gY[i], layers[i]) < tol, i
  assert _rel(G_z, ref) < tol
  logits = eng.discriminate(x.cuda().contiguous(), eng.real)
  want, dl = M.discriminator(P, x, return_layers=True)
  for i in range(4):
    assert _rel(eng.real.Y[i], dl[i]) < tol, i
  assert _rel(logits, want) < 5 * tol     # one 10240-term dot product of the activations above


@pytest.mark.parametrize('math,gtol', [('fp32', 6e-3), ('auto', 6e-2)])
@pytest.mark.parametrize('loss', ['dcgan', 'wgan', 'wgangp'])
def test_train_step_gradients(math, gtol, loss):
  M, P, eng, z, x = _setup(math, 32, 6, loss)
  Pr = {n: t.clone().requires_grad_(True) for n, t in P.items()}
  G_z = M.generator(Pr, z, 32)
  D_x, D_G_z = M.discriminator(Pr, x), M.discriminator(Pr, G_z)
  alpha = torch.rand(6, 1, 1, 1, generator=torch.Generator().manual_seed(8))
  if loss in ('dcgan', 'wgangp'):
    l = M.losses_dim(Pr, z, x, 32, loss, alpha=alpha)
    D_loss, G_loss = l['D_loss'], l['G_loss']
  else:
    D_loss, G_loss = D_G_z.mean() - D_x.mean(), -D_G_z.mean()
  dn, gn = M.d_names(P), M.g_names(P)
  ref_d = dict(zip(dn, torch.autograd.grad(D_loss, [Pr[n] for n in dn], retain_graph=True)))
  ref_g = dict(zip(gn, torch.autograd.grad(G_loss, [Pr[n] for n in gn])))
  # the wgan critic loss is a difference of two means: measure its error against the larger term
  ltol = 2e-3 if math == 'fp32' else 5e-3
  scale_d = max(abs(float(D_loss)), float(D_x.abs().mean()), float(D_G_z.abs().mean()), 1e-2)
  eng.d_step(x.cuda(), z.cuda(), apply=False, alpha=alpha.cuda())
  assert abs(eng.loss_values()[0] - float(D_loss)) < ltol * scale_d
  _check_grads(eng, ref_d, dn, gtol)
  eng.g_step(z.cuda(), apply=False)
  assert abs(eng.loss_values()[1] - float(G_loss)) < ltol * max(abs(float(G_loss)), 1e-2)
  _check_grads(eng, ref_g, gn, gtol)


def test_adam_updates():
  M, P, eng, z, x = _setup('auto', 32, 4)
  before = eng.flat.p.clone()
  eng.train_loop([x.cuda()], [z.cuda()], z.cuda())
  delta = (eng.flat.p - before).abs()
  assert eng.t_d == 1 and eng.t_g == 1 and torch.isfinite(eng.flat.p).all()
  lo, hi = eng.flat.dis_range()
  assert float(delta[lo:hi].max()) > 0 and float(delta[:lo].max()) > 0
  assert float(delta.max()) <= 2.1e-4      # |Adam step| <= lr at t = 1


def test_wgangp_outer_iteration_runs():
  """5 critic steps with the gradient penalty + 1 generator step (train.py:111,149-153)."""
  M, P, eng, z, x = _setup('auto', 32, 4, 'wgangp')
  g = torch.Generator().manual_seed(3)
  xs = [(torch.rand(4, 64, 80, 1, generator=g) * 2 - 1).cuda() for _ in range(5)]
  zs = [torch.randn(4, M.Z_DIM, generator=g).cuda() for _ in range(6)]
  eng.train_loop(xs, zs[:5], zs[5])
  assert eng.t_d == 5 and eng.t_g == 1 and torch.isfinite(eng.flat.p).all()
  d_loss, g_loss = eng.loss_values()
  assert d_loss == d_loss and g_loss == g_loss


@pytest.mark.parametrize('loss', ['dcgan', 'wgangp'])
def test_graph_replay_matches_eager_launches(loss):
  """One D step and one G step replayed from captured CUDA graphs against the same steps launched kernel
  by kernel, from identical state (TF32 path, so the in-place refresh of the packed filters is on the
  replayed path too).  A GAN step at batch 4 amplifies the summation-order noise of the atomically
  accumulated gradients ~50x per iteration (measured: two EAGER engines drift apart just as fast), so
  the engines are re-synchronised after the capturing iteration and the yardstick is the drift
  between two eager engines."""
  from advoc_b200 import _native as N
  from advoc_b200.melspecgan import MelspecGAN
  M, P, eng, z, x = _setup('auto', 32, 4, loss)
  assert eng.use_graphs
  mk = lambda: MelspecGAN({k: v.cuda() for k, v in P.items()}, 4, dim=32, train_loss=loss, math=N.MATH_AUTO,
                          use_graphs=False)
  ref, ref2 = mk(), mk()
  n_d = 5 if loss == 'wgangp' else 1
  g = torch.Generator().manual_seed(6)

  def iteration(it, nd):
    xs = [(torch.rand(4, 64, 80, 1, generator=g) * 2 - 1).cuda() for _ in range(nd)]
    zs = [torch.randn(4, M.Z_DIM, generator=g).cuda() for _ in range(nd + 1)]
    for e in (eng, ref, ref2):
      for k in range(nd):
        alpha = torch.rand(4, 1, 1, 1, generator=torch.Generator().manual_seed(100 * it + k)).cuda()
        e.d_step(xs[k], zs[k], alpha=alpha)
      e.g_step(zs[nd])

  iteration(0, n_d)                             # eager everywhere; `eng` records its graphs
  assert set(eng._graphs) == {'d', 'g', 'adam_D', 'adam_G'} and not ref._graphs
  for e in (eng, ref2):                         # identical state before the replayed iteration
    for name in ('p', 'm', 'v'):
      getattr(e.flat, name).copy_(getattr(ref.flat, name))
    e.load_moving_averages(ref.moving_averages())
    e.refresh_weights()
  p1 = ref.flat.p.clone()
  iteration(1, 1)                               # `eng` replays: one D step + one G step (few steps = little drift)
  assert eng.t_d == ref.t_d == n_d + 1 and eng.t_g == ref.t_g == 2
  # the device-side Adam step sizes follow t although the update graph was captured at t = 1
  for slot, t in ((0, eng.t_d), (1, eng.t_g)):
    want = eng.lr * (1. - eng.b2 ** t) ** 0.5 / (1. - eng.b1 ** t)
    assert abs(float(eng.lr_t[slot]) - want) < 1e-6 * want
  # compare the UPDATE (like test_train_loop_matches_oracle_adam), leaving out the biases whose exact
  # gradient is zero: Adam turns their rounding noise into steps of arbitrary sign
  live = torch.ones_like(p1)
  for n in _DEAD_BIAS:
    o = eng.flat.offsets[n]
    live[o:o + eng.flat.P[n].numel()] = 0
  drift = lambda a, b, f: _rel(f(a) * live, f(b) * live)
  for what, f in (('update', lambda e: e.flat.p - p1), ('m', lambda e: e.flat.m), ('v', lambda e: e.flat.v)):
    noise = drift(ref2, ref, f)
    assert drift(eng, ref, f) < max(5e-3, 5 * noise), (what, noise)
  for n, v in eng.moving_averages().items():
    assert _rel(v, ref.moving_averages()[n]) < 1e-3, n
  a, b = eng.loss_values(), ref.loss_values()
  assert abs(a[0] - b[0]) < 2e-2 * max(1., abs(b[0])) and abs(a[1] - b[1]) < 2e-2 * max(1., abs(b[1]))


@pytest.mark.parametrize('math,tol', [('fp32', 1e-4), ('auto', 1e-3)])
def test_moving_averages_and_inference_graph(math, tol):
  """Two training-mode evaluations advance the batch-norm moving averages like the layer's update ops;
  the training=False graph (models/melspecgan/infer.py:17) then normalises with them."""
  M, P, eng, z, x = _setup(math, 64, 8)
  g = torch.Generator().manual_seed(9)
  z2 = torch.randn(z.shape, generator=g)
  z3 = torch.randn(z.shape, generator=g)
  moving = M.init_moving(64)
  for zi in (z, z2):
    eng.generate(zi.cuda())
    M.generator(P, zi, 64, moving=moving)
  got = eng.moving_averages()
  for n in moving:
    # 0.99 of the value is still the initial 0 / 1: compare the part that moved
    init = 1. if n.endswith('variance') else 0.
    assert _rel(got[n] - init, moving[n] - init) < 10 * tol, n
  # a trained net has moving statistics close to its batch statistics: install those, then infer
  for n, v in moving.items():
    v.copy_(torch.rand(v.shape, generator=g) * 0.05 + (0.02 if n.endswith('variance') else -0.02))
  eng.load_moving_averages({n: v.cuda() for n, v in moving.items()})
  before = {n: v.clone() for n, v in eng.moving_averages().items()}
  G_z = eng.generate(z3.cuda(), training=False)
  ref = M.generator(P, z3, 64, moving=moving, training=False)
  assert _rel(G_z, ref) < tol
  for n, v in eng.moving_averages().items():
    assert torch.equal(v, before[n]), n       # inference leaves them alone


def test_generate_spectrogram_cli_from_tf_checkpoint(tmp_path):
  """scripts/generate_spectrogram.py restores generator weights + moving averages from a TF-1 checkpoint
  prefix (written by the test-side bundle writer) and saves feats_denorm(G(z, training=False))."""
  import os
  import subprocess
  import sys
  import numpy as np
  sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
  import tf_bundle_writer as W
  from oracle import melspecgan_torch as M
  P = M.init_params(seed=3, dim=64)
  g = torch.Generator().manual_seed(4)
  moving = M.init_moving(64)
  momentum, M.BN_MOMENTUM = M.BN_MOMENTUM, 0.   # one evaluation installs that batch's statistics: a "trained" state
  try:
    M.generator(P, torch.randn(16, M.Z_DIM, generator=g), 64, moving=moving)
  finally:
    M.BN_MOMENTUM = momentum
  T = {k: v.numpy() for k, v in P.items() if k.startswith('G/')}
  T.update({k: v.numpy() for k, v in moving.items()})
  T['global_step'] = np.int64(4321)
  prefix = str(tmp_path / 'model.ckpt-4321')
  W.write_bundle(prefix, T, block_size=4096)
  out_dir = str(tmp_path / 'specs')
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([sys.executable, os.path.join(root, 'scripts', 'generate_spectrogram.py'), '--out_dir', out_dir,
                      '--ckpt_fp', prefix, '--n', '6', '--b', '3', '--seed', '7'], capture_output=True, text=True,
                     timeout=600)
  assert r.returncode == 0, r.stderr[-2000:]
  assert 'Restored from step 4321' in r.stdout
  files = sorted(os.listdir(out_dir))
  assert files == ['%09d.npy' % i for i in range(6)]
  gen = torch.Generator(device='cuda')
  gen.manual_seed(7)
  z = torch.randn((3, M.Z_DIM), dtype=torch.float32, device='cuda', generator=gen).cpu()
  ref = (M.generator(P, z, 64, moving=moving, training=False) + 1.) * 0.5
  for j in range(3):
    s = np.load(os.path.join(out_dir, files[j]))
    assert s.shape == (64, 80, 1) and s.dtype == np.float32
    assert _rel(torch.from_numpy(s), ref[j]) < 1e-3
