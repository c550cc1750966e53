"""GPU parity of the conv stacks (through the C-ABI) against the PyTorch-CPU oracle
(oracle/nets_torch.py; parity unpinned by the reference -- see its header).
Tolerance: 1e-3 relative L2 (BASELINE.json north_star) for the TF32 tensor-core path; the
exact-fp32 CUDA-core path is held to 1e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-3
TOL_FP32 = 1e-5


def _rel(a, b):
  a = a.detach().double().cpu()
  b = b.detach().double().cpu()
  return float((a - b).norm() / b.norm().clamp_min(1e-300))


def _to_cuda(P):
  return {k: v.cuda().contiguous() for k, v in P.items()}


def _tf32(x):
  """Round to TF32 (nearest, ties away) like the producing layer's epilogue does."""
  i = x.contiguous().view(torch.int32)
  return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def _weights_for(L, kd, ldx):
  from advoc_b200 import nets
  wp = nets._pack_for_tc(L, kd, ldx)
  return kd if wp is None else wp


def _randomize_biases(P, seed=11):
  g = torch.Generator().manual_seed(seed)
  for k in P:
    if k.endswith('/bias'):
      P[k] = torch.randn(P[k].shape, generator=g) * 0.05
  return P


@pytest.mark.parametrize('math', ['fp32', 'auto'])
def test_single_conv_layers(math):
  """conv k4 s2 SAME on odd width, PatchGAN pad-1 VALID s1/s2, deconv with crop + dual write."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  mm = N.MATH_FP32 if math == 'fp32' else N.MATH_AUTO
  tol = TOL_FP32 if math == 'fp32' else TOL
  g = torch.Generator().manual_seed(0)
  B = 2
  for (H, W, Cin, Cout, sh) in [(16, 33, 64, 128, 2), (8, 17, 32, 64, 2), (1, 5, 64, 32, 1)]:
    x = _tf32(torch.randn(B, H, W, Cin, generator=g))
    k = torch.randn(4, 4, Cin, Cout, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    ref = O.conv_same(x, k, b, (sh, 2))
    ho, pt, _ = nets.same_pads(H, 4, sh)
    wo, pl, _ = nets.same_pads(W, 4, 2)
    y0 = torch.full((B, ho, wo, Cout), float('nan'), device='cuda')
    cat = torch.full((B, ho, wo, Cout + 32), float('nan'), device='cuda')
    L = nets._Conv('t', 'conv', nets._desc(B, H, W, Cin, Cout, sh, 2, pt, pl, ho, wo, mm))
    bd, xd, kd = b.cuda(), x.cuda(), k.cuda()   # keep alive: the epilogue holds raw pointers
    ep = nets._epilogue(bd, y0, Cout, 0, N.ACT_LRELU, cat, Cout + 32, 32, N.ACT_RELU)
    L.run(xd, Cin, _weights_for(L, kd, Cin), ep)
    torch.cuda.synchronize()
    assert N.debug_flags() == 0
    assert _rel(y0, O.lrelu(ref)) < tol
    assert _rel(cat[..., 32:], torch.relu(ref)) < tol
    assert torch.isnan(cat[..., :32]).all()
  # discriminator conv: explicit pad 1, VALID, strides 2 and 1
  for (H, W, Cin, Cout, st) in [(32, 65, 32, 64, 2), (16, 32, 64, 96, 1)]:
    x = _tf32(torch.randn(B, H, W, Cin, generator=g))
    k = torch.randn(4, 4, Cin, Cout, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    ref = torch.sigmoid(O.discrim_conv(x, k, b, st))
    ho, wo = (H + 2 - 4) // st + 1, (W + 2 - 4) // st + 1
    y = torch.empty((B, ho, wo, Cout), device='cuda')
    L = nets._Conv('t', 'conv', nets._desc(B, H, W, Cin, Cout, st, st, 1, 1, ho, wo, mm))
    bd, xd, kd = b.cuda(), x.cuda(), k.cuda()
    L.run(xd, Cin, _weights_for(L, kd, Cin), nets._epilogue(bd, y, Cout, 0, N.ACT_SIGMOID))
    assert N.debug_flags() == 0
    assert _rel(y, ref) < tol
  # deconv k4 s2 SAME, last column cropped, relu, written at channel offset 0 of a wider buffer
  for (H, W, Cin, Cout, sh) in [(8, 17, 64, 32, 2), (1, 3, 32, 64, 1)]:
    x = _tf32(torch.relu(torch.randn(B, H, W, Cin, generator=g)))
    k = torch.randn(4, 4, Cout, Cin, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    ref = torch.relu(O.deconv_same(x, k, b, (sh, 2)))[:, :, :-1, :]
    cat = torch.full((B, H * sh, 2 * W - 1, Cout + 16), float('nan'), device='cuda')
    L = nets._Conv('t', 'deconv', nets._desc(B, H * sh, 2 * W, Cout, Cin, sh, 2, 1, 1, H, W, mm))
    bd, xd, kd = b.cuda(), x.cuda(), k.cuda()
    ep = nets._epilogue(bd, cat, Cout + 16, 0, N.ACT_RELU, store_w=2 * W - 1)
    L.run(xd, Cin, _weights_for(L, kd, Cin), ep)
    assert N.debug_flags() == 0
    assert _rel(cat[..., :Cout], ref) < tol
    assert torch.isnan(cat[..., Cout:]).all()


@pytest.mark.parametrize('math,tol', [('fp32', TOL_FP32), ('auto', TOL)])
def test_small_generator_forward_matches_oracle(math, tol):
  """BASELINE configs[1] net (AdVoc-small) at batch 2, dropout off, every stored activation."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from advoc_b200.model import SpectralUtil
  from oracle import nets_torch as O
  P = _randomize_biases(O.init_params(O.SMALL, seed=0))
  g = torch.Generator().manual_seed(1)
  mel = torch.randn(2, 256, 80, 1, generator=g).abs()
  su = SpectralUtil()
  x = su.mel_linear_to_mag_spec(mel.cuda())
  spec = nets.GenSpec(32, 5, (5, 4))
  G = nets.Generator(spec, _to_cuda(P), 2, N.MATH_FP32 if math == 'fp32' else N.MATH_AUTO)
  out = G.forward(x.contiguous())
  ref, layers = O.generator(P, x.cpu(), O.SMALL, return_layers=True)
  assert out.shape == ref.shape == (2, 256, 513, 1)
  assert _rel(out, ref) < tol
  # intermediate buffers: lrelu(encoder_i) and the relu'ed decoder concat inputs
  for i in range(1, 5):
    assert _rel(G.E[i], O.lrelu(layers[i - 1])) < tol
  assert _rel(G.Cat[5], torch.relu(layers[4])) < tol
  cat4 = torch.relu(torch.cat([layers[5][:, :, :-1, :], layers[3]], dim=3))
  assert _rel(G.Cat[4], cat4) < tol


def test_generator_dropout_masks_injected():
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  P = O.init_params(O.SMALL, seed=3)
  g = torch.Generator().manual_seed(2)
  x = torch.randn(1, 256, 513, 1, generator=g).abs() * 0.1
  spec = nets.GenSpec(32, 5, (5, 4))
  G = nets.Generator(spec, _to_cuda(P), 1, N.MATH_FP32)
  masks = {k: (torch.rand(G.dropout_shape(k), generator=g) < 0.5) for k in (5, 4)}
  # the oracle masks the un-cropped decoder output; pad the cropped column back with zeros
  full = {k: torch.cat([m, torch.zeros(m.shape[0], m.shape[1], 1, m.shape[3], dtype=torch.bool)], 2)
          for k, m in masks.items()}
  ref = O.generator(P, x, O.SMALL, {k: m.float() for k, m in full.items()})
  out = G.forward(x.cuda(), dropout={k: m.to(torch.uint8).cuda().contiguous()
                                     for k, m in masks.items()})
  assert _rel(out, ref) < TOL_FP32
  # rng mode: deterministic per seed, different across seeds, roughly half dropped
  a = G.forward(x.cuda(), dropout='rng', seed=5).clone()
  b = G.forward(x.cuda(), dropout='rng', seed=5).clone()
  c = G.forward(x.cuda(), dropout='rng', seed=6).clone()
  assert torch.equal(a, b) and not torch.equal(a, c)
  frac = float((G.Cat[4][..., :256] == 0).float().mean())
  assert 0.45 < frac < 0.95


@pytest.mark.parametrize('math,tol', [('fp32', TOL_FP32), ('auto', TOL)])
def test_discriminator_forward_matches_oracle(math, tol):
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  P = _randomize_biases(O.init_params(O.SMALL, seed=0))
  g = torch.Generator().manual_seed(4)
  a = torch.randn(2, 256, 513, 1, generator=g).abs()
  b = torch.randn(2, 256, 513, 1, generator=g).abs()
  D = nets.Discriminator(32, _to_cuda(P), 2, N.MATH_FP32 if math == 'fp32' else N.MATH_AUTO)
  out = D.forward(torch.cat([a, b], 3).cuda().contiguous())
  ref, layers = O.discriminator(P, a, b, return_layers=True)
  assert out.shape == ref.shape == (2, 30, 62, 1)
  for got, want in zip(D.act, layers):
    assert _rel(got, want) < tol


def test_model_api_mirror():
  from advoc_b200.model import Advoc, AdvocSmall, Modes, override_model_attrs
  m = AdvocSmall(Modes.INFER)
  m, summary = override_model_attrs(m, 'subseq_len=256,gan_weight=0.5')
  assert m.gan_weight == 0.5 and 'ngf,32' in summary
  x = torch.rand(1, 256, 513, 1, device='cuda')
  y = m.build_generator(x, dropout=None)
  assert y.shape == (1, 256, 513, 1) and torch.isfinite(y).all()
  p = m.build_discriminator(x, y)
  assert p.shape == (1, 30, 62, 1) and float(p.min()) > 0 and float(p.max()) < 1
  assert Advoc.ngf == 64 and Advoc.train_batch_size == 8


# ---------------------------------------------------------------------------------------------
# round 2: the configurations bench.py actually times (B = 32 small, regular model, wide N tiles)
# and the fp16-operand inference mode
# ---------------------------------------------------------------------------------------------
def _check_generator(G, P, x, spec_o, tol, n_enc):
  """Every stored activation of a forward pass against the oracle's layer list."""
  from oracle import nets_torch as O
  out = G.forward(x.contiguous())
  torch.cuda.synchronize()
  from advoc_b200 import _native as N
  assert N.debug_flags() == 0
  ref, layers = O.generator(P, x.cpu(), spec_o, return_layers=True)
  errs = {'out': _rel(out, ref)}
  for i in range(1, n_enc):
    want = O.lrelu(layers[i - 1])
    errs['E%d' % i] = _rel(G.E[i][:, :, :want.shape[2]], want)   # (the fp16 small generator pads E1's rows)
  errs['Cat%d' % n_enc] = _rel(G.Cat[n_enc], torch.relu(layers[n_enc - 1]))
  for k in range(n_enc - 1, 0, -1):
    dec = layers[n_enc + (n_enc - 1 - k)]          # output of decoder_{k+1}
    cat = torch.relu(torch.cat([dec[:, :, :-1, :], layers[k - 1]], dim=3))
    errs['Cat%d' % k] = _rel(G.Cat[k], cat)
  bad = {k: v for k, v in errs.items() if not v < tol}
  assert not bad, (bad, errs)
  return errs


def _instantiations(G):
  return sorted({(L.kernel_family().split('/')[0], L.tile_n(), L.half_operands())
                 for L in list(G.enc.values()) + list(G.dec.values())})


@pytest.mark.parametrize('math', ['auto', 'f16'])
def test_small_generator_forward_at_the_benchmarked_batch(math):
  """BASELINE configs[1] exactly as bench.py times it (AdVoc-small, B = 32): every stored activation
  within 1e-3 of the oracle, and the tile instantiations of that batch size really ran
  (conv_tc_kernel<128,3> only appears from B = 32 on)."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from advoc_b200.model import SpectralUtil
  from oracle import nets_torch as O
  P = _randomize_biases(O.init_params(O.SMALL, seed=0))
  g = torch.Generator().manual_seed(1)
  mel = torch.randn(32, 256, 80, 1, generator=g).abs()
  x = SpectralUtil().mel_linear_to_mag_spec(mel.cuda())
  G = nets.Generator(nets.GenSpec(32, 5, (5, 4)), _to_cuda(P), 32, N.MATH_F16 if math == 'f16' else N.MATH_AUTO)
  _check_generator(G, P, x, O.SMALL, TOL, 5)
  inst = _instantiations(G)
  fams = {k for k, _, _ in inst}
  import os
  if os.environ.get('ADVOC_NO_P2D') or os.environ.get('ADVOC_P2D_FORCE') or os.environ.get('ADVOC_NO_ONE_IN_TC'):
    return          # forced kernel choice (tests/test_gpu_forced_paths.py): the routing below does not apply
  assert 'conv_tc_kernel' in fams and 'deconv_one_tc_kernel' in fams and 'conv_one_in_tc_kernel' in fams, inst
  if math == 'auto':
    assert ('conv_tc_kernel', 128, False) in inst, inst
  else:
    assert any(h for _, _, h in inst), inst
    assert G.Cat[1].dtype == torch.float16 and G.Cat[4].dtype == torch.float16
    assert G.pair2 is not None and G.E[1].dtype == torch.float16 and bool((G.E[1][:, :, -1] == 0).all())


@pytest.mark.parametrize('math', ['auto', 'f16'])
def test_regular_generator_forward_matches_oracle(math):
  """AdVoc regular (ngf 64, 8 + 8 layers, the 1x3 bottleneck and the 1024-channel concats of BASELINE
  configs[2]/[3]) at B = 4, every stored activation."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  P = _randomize_biases(O.init_params(O.REGULAR, seed=0))
  g = torch.Generator().manual_seed(2)
  x = torch.randn(4, 256, 513, 1, generator=g).abs() * 0.5
  G = nets.Generator(nets.GenSpec(64, 8, (8, 7, 6)), _to_cuda(P), 4, N.MATH_F16 if math == 'f16' else N.MATH_AUTO)
  _check_generator(G, P, x.cuda(), O.REGULAR, TOL, 8)


def test_regular_discriminator_forward_matches_oracle():
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  P = _randomize_biases(O.init_params(O.REGULAR, seed=0))
  g = torch.Generator().manual_seed(4)
  a = torch.randn(4, 256, 513, 1, generator=g).abs()
  b = torch.randn(4, 256, 513, 1, generator=g).abs()
  D = nets.Discriminator(64, _to_cuda(P), 4, N.MATH_AUTO)
  out = D.forward(torch.cat([a, b], 3).cuda().contiguous())
  ref, layers = O.discriminator(P, a, b, return_layers=True)
  assert N.debug_flags() == 0
  for got, want in zip(D.act, layers):
    assert _rel(got, want) < TOL


def test_wide_n_tiles_regular_model():
  """conv_tc_kernel<256,4> (19.5 % of the regular train step at B = 32) is only picked by itself from
  large batches on; ADVOC_TC_WIDE_N=1 makes the per-tap kernel always take its widest tile, so a child
  process covers that instantiation with the regular generator / discriminator forward at B = 4."""
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  env = dict(os.environ)
  env['ADVOC_TC_WIDE_N'] = '1'
  env['ADVOC_EXPECT_WIDE_N'] = '1'
  r = subprocess.run([sys.executable, '-m', 'pytest', '-x', '-q', '-m', 'gpu', 'tests/test_gpu_nets.py', '-k',
                      'regular_generator_forward or regular_discriminator or wide_n_probe'],
                     cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
  assert r.returncode == 0, r.stdout[-3000:]


def test_wide_n_probe():
  """In the ADVOC_TC_WIDE_N child: the regular stacks contain 256-wide per-tap launches."""
  import os
  if not os.environ.get('ADVOC_EXPECT_WIDE_N'):
    pytest.skip('runs inside test_wide_n_tiles_regular_model')
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  P = O.init_params(O.REGULAR, seed=0)
  G = nets.Generator(nets.GenSpec(64, 8, (8, 7, 6)), _to_cuda(P), 4, N.MATH_AUTO)
  G.forward(torch.rand(4, 256, 513, 1, device='cuda'))
  assert ('conv_tc_kernel', 256, False) in _instantiations(G), _instantiations(G)


@pytest.mark.parametrize('kernel', ['auto', 'p2d', 'tc'])
def test_single_layers_fp16_operands(kernel):
  """ADVOC_MATH_F16: fp16 activations and filters in, fp16 (dual-write, channel offset, crop) or fp32
  out, on both tcgen05 kernels (forced through ADVOC_P2D_FORCE / ADVOC_NO_P2D in child processes)."""
  import os
  import subprocess
  import sys
  if kernel != 'auto':
    if os.environ.get('ADVOC_F16_CHILD'):
      pytest.skip('child')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    env['ADVOC_P2D_FORCE' if kernel == 'p2d' else 'ADVOC_NO_P2D'] = '1'
    env['ADVOC_F16_CHILD'] = '1'
    r = subprocess.run([sys.executable, '-m', 'pytest', '-x', '-q', '-m', 'gpu', 'tests/test_gpu_nets.py', '-k',
                        'single_layers_fp16_operands and auto'],
                       cwd=root, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    assert r.returncode == 0, r.stdout[-3000:]
    return
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  g = torch.Generator().manual_seed(0)
  B = 2
  q = lambda t: t.half().float()
  for (H, W, Cin, Cout, sh, odt) in [(16, 33, 64, 128, 2, torch.float16), (32, 65, 128, 64, 2, torch.float16),
                                      (8, 17, 64, 32, 2, torch.float16), (1, 5, 64, 32, 1, torch.float32)]:
    x = torch.randn(B, H, W, Cin, generator=g)
    k = torch.randn(4, 4, Cin, Cout, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    ref = O.conv_same(q(x), q(k), b, (sh, 2))
    ho, pt, _ = nets.same_pads(H, 4, sh)
    wo, pl, _ = nets.same_pads(W, 4, 2)
    y0 = torch.full((B, ho, wo, Cout), float('nan'), device='cuda', dtype=odt)
    cat = torch.full((B, ho, wo, Cout + 64), float('nan'), device='cuda', dtype=odt)
    L = nets._Conv('t', 'conv', nets._desc(B, H, W, Cin, Cout, sh, 2, pt, pl, ho, wo, N.MATH_F16))
    assert L.path(Cin) == N.MATH_F16
    bd, xd, kd = b.cuda(), x.cuda().half(), k.cuda()
    wp = nets._pack_for_tc(L, kd, Cin)
    assert wp.dtype == torch.float16
    ep = nets._epilogue(bd, y0, Cout, 0, N.ACT_LRELU, cat, Cout + 64, 64, N.ACT_RELU)
    L.run(xd, Cin, wp, ep)
    torch.cuda.synchronize()
    assert N.debug_flags() == 0
    assert _rel(y0.float(), O.lrelu(ref)) < TOL, (H, W, Cin, Cout)
    assert _rel(cat[..., 64:].float(), torch.relu(ref)) < TOL
    assert torch.isnan(cat[..., :64]).all()
  for (H, W, Cin, Cout, sh, odt) in [(8, 17, 64, 32, 2, torch.float16), (16, 33, 128, 64, 2, torch.float16),
                                      (1, 3, 64, 64, 1, torch.float16), (8, 17, 192, 128, 2, torch.float32)]:
    x = torch.relu(torch.randn(B, H, W, Cin, generator=g))
    k = torch.randn(4, 4, Cout, Cin, generator=g) * 0.05
    b = torch.randn(Cout, generator=g) * 0.1
    ref = torch.relu(O.deconv_same(q(x), q(k), b, (sh, 2)))[:, :, :-1, :]
    cat = torch.full((B, H * sh, 2 * W - 1, Cout + 32), float('nan'), device='cuda', dtype=odt)
    L = nets._Conv('t', 'deconv', nets._desc(B, H * sh, 2 * W, Cout, Cin, sh, 2, 1, 1, H, W, N.MATH_F16))
    assert L.path(Cin) == N.MATH_F16
    bd, xd, kd = b.cuda(), x.cuda().half(), k.cuda()
    ep = nets._epilogue(bd, cat, Cout + 32, 0, N.ACT_RELU, store_w=2 * W - 1)
    L.run(xd, Cin, nets._pack_for_tc(L, kd, Cin), ep)
    torch.cuda.synchronize()
    assert N.debug_flags() == 0
    assert _rel(cat[..., :Cout].float(), ref) < TOL, (H, W, Cin, Cout)
    assert torch.isnan(cat[..., Cout:]).all()
  # decoder_1 geometry: 64 fp16 channels -> one fp32 channel, last column cropped
  H, W, Cin = 16, 33, 64
  x = torch.relu(torch.randn(B, H, W, Cin, generator=g))
  k = torch.randn(4, 4, 1, Cin, generator=g) * 0.05
  b = torch.randn(1, generator=g) * 0.1
  ref = O.deconv_same(q(x), q(k), b, (2, 2))[:, :, :-1, :]
  out = torch.full((B, 2 * H, 2 * W - 1, 1), float('nan'), device='cuda')
  L = nets._Conv('t', 'deconv', nets._desc(B, 2 * H, 2 * W, 1, Cin, 2, 2, 1, 1, H, W, N.MATH_F16))
  assert L.path(Cin) == N.MATH_F16
  bd, xd, kd = b.cuda(), x.cuda().half(), k.cuda()
  L.run(xd, Cin, nets._pack_for_tc(L, kd, Cin), nets._epilogue(bd, out, 1, 0, N.ACT_NONE, store_w=2 * W - 1))
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert L.kernel_family().startswith('deconv_one_tc')
  assert _rel(out, ref) < TOL


@pytest.mark.parametrize('cout', [32, 64, 128])
def test_one_input_channel_conv_on_tensor_cores(cout):
  """encoder_1's geometry (one input channel, k4 s2 SAME, odd width) on conv_one_in_tc_kernel: the 3 x tf32
  split keeps the products fp32-exact (<= 2e-6 against the oracle, where a single tf32 pass would be
  3e-4), with fp32 + fp32, fp16 + fp16 (padded rows) and single-output epilogues."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  g = torch.Generator().manual_seed(cout)
  B, H, W = 3, 20, 37
  x = torch.randn(B, H, W, 1, generator=g).abs() * 3.0
  k = torch.randn(4, 4, 1, cout, generator=g) * 0.05
  b = torch.randn(cout, generator=g) * 0.1
  ref = O.conv_same(x, k, b, (2, 2))
  ho, pt, _ = nets.same_pads(H, 4, 2)
  wo, pl, _ = nets.same_pads(W, 4, 2)
  L = nets._Conv('t', 'conv', nets._desc(B, H, W, 1, cout, 2, 2, pt, pl, ho, wo, N.MATH_AUTO))
  xd, kd, bd = x.cuda(), k.cuda(), b.cuda()
  # fp32 dual write into a wider buffer at a channel offset
  y0 = torch.full((B, ho, wo, cout), float('nan'), device='cuda')
  cat = torch.full((B, ho, wo, cout + 64), float('nan'), device='cuda')
  L.run(xd, 1, kd, nets._epilogue(bd, y0, cout, 0, N.ACT_LRELU, cat, cout + 64, 64, N.ACT_RELU))
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert L.kernel_family().startswith('conv_one_in_tc')
  assert _rel(y0, O.lrelu(ref)) < 2e-6 and _rel(cat[..., 64:], torch.relu(ref)) < 2e-6
  assert torch.isnan(cat[..., :64]).all()
  # fp16 outputs, out0 with one padded pixel per row
  h0 = torch.zeros((B, ho, wo + 1, cout), device='cuda', dtype=torch.float16)
  h1 = torch.full((B, ho, wo, cout + 8), float('nan'), device='cuda', dtype=torch.float16)
  L.run(xd, 1, kd, nets._epilogue(bd, h0, cout, 0, N.ACT_LRELU, h1, cout + 8, 8, N.ACT_RELU, row_pad0=1))
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert _rel(h0[:, :, :wo].float(), O.lrelu(ref)) < 5e-4 and bool((h0[:, :, wo] == 0).all())
  assert _rel(h1[..., 8:].float(), torch.relu(ref)) < 5e-4
  # single fp32 output, no activation, no bias
  y2 = torch.empty((B, ho, wo, cout), device='cuda')
  L.run(xd, 1, kd, nets._epilogue(None, y2, cout, 0, N.ACT_NONE))
  torch.cuda.synchronize()
  assert _rel(y2, ref - b) < 2e-6


@pytest.mark.parametrize('cout', [32, 64])
def test_two_input_channel_conv_on_tensor_cores(cout):
  """Discriminator layer_1's geometry (two input channels, pad 1 + VALID k4 s2, odd width) on
  conv_one_in_tc_kernel<.., MODE 1>: K = 32, hi / lo parts in two A tiles, fp32-exact products."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  g = torch.Generator().manual_seed(100 + cout)
  B, H, W = 3, 22, 45
  x = torch.randn(B, H, W, 2, generator=g) * 2.0
  k = torch.randn(4, 4, 2, cout, generator=g) * 0.05
  b = torch.randn(cout, generator=g) * 0.1
  ref = O.discrim_conv(x, k, b, 2)
  ho, wo = (H + 2 - 4) // 2 + 1, (W + 2 - 4) // 2 + 1
  assert tuple(ref.shape) == (B, ho, wo, cout)
  L = nets._Conv('t', 'conv', nets._desc(B, H, W, 2, cout, 2, 2, 1, 1, ho, wo, N.MATH_AUTO))
  xd, kd, bd = x.cuda(), k.cuda(), b.cuda()
  y = torch.full((B, ho, wo, cout), float('nan'), device='cuda')
  L.run(xd, 2, kd, nets._epilogue(bd, y, cout, 0, N.ACT_LRELU))
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert L.kernel_family().startswith('conv_one_in_tc')
  assert _rel(y, O.lrelu(ref)) < 2e-6
  # fp16 destination with a channel offset
  h = torch.full((B, ho, wo, cout + 8), float('nan'), device='cuda', dtype=torch.float16)
  L.run(xd, 2, kd, nets._epilogue(bd, h, cout + 8, 8, N.ACT_RELU))
  torch.cuda.synchronize()
  assert _rel(h[..., 8:].float(), torch.relu(ref)) < 5e-4


@pytest.mark.parametrize('cout,split', [(128, 64), (64, 0), (256, 128)])
def test_one_input_channel_conv_backward_epilogue(cout, split):
  """decoder_1's input gradient: a conv from the one-channel d loss / d generated to the concat gradient, gated
  by relu' of the stored concat buffer with one scale per half (dropout 1 / keep on the decoder half);
  Cout = 256 runs as two blockIdx.y chunks."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  g = torch.Generator().manual_seed(7 + cout)
  B, H, W = 2, 36, 50
  x = torch.randn(B, H, W, 1, generator=g)
  k = torch.randn(4, 4, 1, cout, generator=g) * 0.05
  gate = torch.randn(B, H // 2, W // 2, cout + 8, generator=g)
  ref = O.conv_same(x, k, torch.zeros(cout), (2, 2))
  ho, pt, _ = nets.same_pads(H, 4, 2)
  wo, pl, _ = nets.same_pads(W, 4, 2)
  gs = gate[..., 8:]
  scale = torch.where(torch.arange(cout) < split, torch.tensor(2.0), torch.tensor(0.5))
  xd, kd, gd = x.cuda(), k.cuda(), gate.cuda()     # (kept alive: the epilogue holds raw pointers)
  for act, neg in ((N.ACT_RELU, 0.0), (N.ACT_LRELU, 0.2)):
    want = ref * torch.where(gs > 0, torch.tensor(1.0), torch.tensor(neg)) * scale
    L = nets._Conv('t', 'conv', nets._desc(B, H, W, 1, cout, 2, 2, pt, pl, ho, wo, N.MATH_AUTO))
    y = torch.full((B, ho, wo, cout + 16), float('nan'), device='cuda')
    ep = nets._epilogue(None, y, cout + 16, 16, N.ACT_NONE, gate=gd, ld_gate=cout + 8, coff_gate=8,
                        gate_act=act, gate_split=split, gate_scale0=2.0, gate_scale1=0.5)
    L.run(xd, 1, kd, ep)
    torch.cuda.synchronize()
    assert N.debug_flags() == 0
    assert _rel(y[..., 16:], want) < 2e-6
    assert torch.isnan(y[..., :16]).all()


@pytest.mark.parametrize('cin', [64, 256, 512])
def test_transposed_conv_from_one_channel_on_tensor_cores(cin):
  """Input gradient of the PatchGAN head (k4 s1, pad 1 + VALID, Cin -> 1): a stride-1 transposed conv from one
  channel, gated by lrelu' of the stored layer_4 activation; checked against autograd of the oracle's conv."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  g = torch.Generator().manual_seed(cin)
  B, H, W = 2, 13, 21
  a = torch.randn(B, H, W, cin, generator=g)                 # stored (post-lrelu) activation of layer_4
  k = torch.randn(4, 4, cin, 1, generator=g) * 0.05
  xin = a.clone().requires_grad_(True)
  out = O.discrim_conv(xin, k, torch.zeros(1), 1)
  ho, wo = H + 2 - 4 + 1, W + 2 - 4 + 1
  assert tuple(out.shape) == (B, ho, wo, 1)
  dy = torch.randn(B, ho, wo, 1, generator=g)
  out.backward(dy)
  want = xin.grad * torch.where(a > 0, torch.tensor(1.0), torch.tensor(0.2))
  L = nets._Conv('t', 'deconv', nets._desc(B, H, W, cin, 1, 1, 1, 1, 1, ho, wo, N.MATH_AUTO))
  y = torch.full((B, H, W, cin), float('nan'), device='cuda')
  ad, dyd, kd = a.cuda(), dy.cuda(), k.cuda()      # (kept alive: the epilogue holds raw pointers)
  ep = nets._epilogue(None, y, cin, 0, N.ACT_NONE, gate=ad, ld_gate=cin, gate_act=N.ACT_LRELU, round_tf32=0)
  L.run(dyd, 1, kd, ep)
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert _rel(y, want) < 2e-6


@pytest.mark.parametrize('math', ['auto', 'f16'])
def test_bottleneck_layers_split_k(math):
  """The regular model's bottleneck geometry (a handful of output tiles, 512-channel filters) on the per-tap
  kernel with its tap / channel loop split over CTAs: partial accumulators to the workspace, then
  splitk_finalize_kernel = ordered sum + the layer's epilogue (bias, activation, dual write, crop)."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  half = math == 'f16'
  m = N.MATH_F16 if half else N.MATH_AUTO
  q = (lambda t: t.half().float()) if half else _tf32
  adt = torch.float16 if half else torch.float32
  g = torch.Generator().manual_seed(3)
  B, H, W, Cin, Cout = 4, 2, 5, 512, 256
  x = q(torch.randn(B, H, W, Cin, generator=g))       # exactly representable operands
  k = torch.randn(4, 4, Cin, Cout, generator=g) * 0.02
  b = torch.randn(Cout, generator=g) * 0.1
  ref = O.conv_same(q(x), q(k), b, (2, 2))
  ho, pt, _ = nets.same_pads(H, 4, 2)
  wo, pl, _ = nets.same_pads(W, 4, 2)
  L = nets._Conv('t', 'conv', nets._desc(B, H, W, Cin, Cout, 2, 2, pt, pl, ho, wo, m))
  L.ldx = Cin
  assert L.kernel_family().startswith('conv_tc')
  y0 = torch.full((B, ho, wo, Cout), float('nan'), device='cuda', dtype=adt)
  cat = torch.full((B, ho, wo, Cout + 64), float('nan'), device='cuda', dtype=adt)
  bd, xd, kd = b.cuda(), x.cuda().to(adt), k.cuda()
  wp = nets._pack_for_tc(L, kd, Cin)
  ep = nets._epilogue(bd, y0, Cout, 0, N.ACT_LRELU, cat, Cout + 64, 64, N.ACT_RELU)
  n0 = N.launch_count()
  L.run(xd, Cin, wp, ep)
  assert N.launch_count() - n0 == 2            # split-K kernel + finalize
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert _rel(y0.float(), O.lrelu(ref)) < TOL
  assert _rel(cat[..., 64:].float(), torch.relu(ref)) < TOL
  assert torch.isnan(cat[..., :64].float()).all()
  # transposed, cropped: decoder_8-like
  H, W, Cin, Cout = 1, 3, 512, 256
  x = q(torch.relu(torch.randn(B, H, W, Cin, generator=g)))
  k = torch.randn(4, 4, Cout, Cin, generator=g) * 0.02
  ref = torch.relu(O.deconv_same(q(x), q(k), b, (2, 2)))[:, :, :-1, :]
  out = torch.full((B, 2 * H, 2 * W - 1, Cout + 32), float('nan'), device='cuda', dtype=adt)
  L = nets._Conv('t', 'deconv', nets._desc(B, 2 * H, 2 * W, Cout, Cin, 2, 2, 1, 1, H, W, m))
  L.ldx = Cin
  assert L.kernel_family().startswith('conv_tc')
  xd, kd = x.cuda().to(adt), k.cuda()
  ep = nets._epilogue(bd, out, Cout + 32, 0, N.ACT_RELU, store_w=2 * W - 1)
  wp = nets._pack_for_tc(L, kd, Cin)
  n0 = N.launch_count()
  L.run(xd, Cin, wp, ep)
  assert N.launch_count() - n0 == 2
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert _rel(out[..., :Cout].float(), ref) < TOL
  assert torch.isnan(out[..., Cout:].float()).all()


@pytest.mark.parametrize('cin', [256, 512])
def test_conv_to_one_channel_on_tensor_cores(cin):
  """PatchGAN head (k4 s1, pad 1 + VALID, Cin -> 1, bias + sigmoid) as a 1x1 tensor-core convolution to the 16
  tap sums per input pixel + a gather (conv_tc.cu: conv_to_one_tc), fp32-accurate on
  tf32-representable operands."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  g = torch.Generator().manual_seed(cin + 1)
  B, H, W = 3, 13, 21
  x = _tf32(torch.randn(B, H, W, cin, generator=g))
  k = _tf32(torch.randn(4, 4, cin, 1, generator=g) * 0.05)
  b = torch.randn(1, generator=g) * 0.1
  ref = torch.sigmoid(O.discrim_conv(x, k, b, 1))
  ho, wo = H + 2 - 4 + 1, W + 2 - 4 + 1
  L = nets._Conv('t', 'conv', nets._desc(B, H, W, cin, 1, 1, 1, 1, 1, ho, wo, N.MATH_AUTO))
  xd, kd, bd = x.cuda(), k.cuda(), b.cuda()
  y = torch.full((B, ho, wo, 1), float('nan'), device='cuda')
  for _ in range(2):                  # the second call reuses the scratch
    n0 = N.launch_count()
    L.run(xd, cin, kd, nets._epilogue(bd, y, 1, 0, N.ACT_SIGMOID))
    assert N.launch_count() - n0 in (3, 4)     # pack, 1x1 conv (+ split-K finalize at this tiny size), gather
    torch.cuda.synchronize()
    assert N.debug_flags() == 0
    assert _rel(y, ref) < 1e-5


@pytest.mark.parametrize('cout', [64, 128])
def test_5x5_one_input_channel_conv_on_tensor_cores(cout):
  """MelspecGAN conv_0's geometry (5x5 s2 SAME from one channel, models/melspecgan/conv2d.py:182-184) on
  conv_one_in_tc_kernel<.., MODE 3>: 25 taps padded to K = 32; plain forward and the gated epilogue the
  gradient-penalty up-sweep uses."""
  import torch.nn.functional as F
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  g = torch.Generator().manual_seed(500 + cout)
  B, H, W = 3, 28, 41
  x = torch.randn(B, H, W, 1, generator=g)
  k = torch.randn(5, 5, 1, cout, generator=g) * 0.05
  b = torch.randn(cout, generator=g) * 0.1
  ho, pt, pb = nets.same_pads(H, 5, 2)
  wo, pl, pr = nets.same_pads(W, 5, 2)
  xt = F.pad(x.permute(0, 3, 1, 2).double(), (pl, pr, pt, pb))
  ref = F.conv2d(xt, k.double().permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1).float()
  assert tuple(ref.shape) == (B, ho, wo, cout)
  L = nets._Conv('t', 'conv', N.ConvDesc(B, H, W, 1, cout, 5, 5, 2, 2, pt, pl, ho, wo, N.MATH_AUTO))
  xd, kd, bd = x.cuda(), k.cuda(), b.cuda()
  y = torch.full((B, ho, wo, cout), float('nan'), device='cuda')
  L.run(xd, 1, kd, nets._epilogue(bd, y, cout, 0, N.ACT_LRELU))
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert L.kernel_family().startswith('conv_one_in_tc')
  refb = ref + b
  assert _rel(y, torch.where(refb > 0, refb, 0.2 * refb)) < 2e-6
  gate = torch.randn(B, ho, wo, cout, generator=g)
  gd = gate.cuda()
  y2 = torch.full((B, ho, wo, cout), float('nan'), device='cuda')
  L.run(xd, 1, kd, nets._epilogue(None, y2, cout, 0, N.ACT_NONE, gate=gd, ld_gate=cout, gate_act=N.ACT_LRELU))
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert _rel(y2, ref * torch.where(gate > 0, torch.tensor(1.0), torch.tensor(0.2))) < 2e-6


@pytest.mark.parametrize('math,cout', [('f16', 32), ('f16', 64), ('auto', 32), ('auto', 64)])
def test_transposed_conv_merged_parity_classes(math, cout):
  """conv_tc_merged_kernel: k4 s2 transposed convolution with the four output parity classes computed from nine
  shared shifted views of the input tile (decoder_2 / decoder_3 geometry: many positions, 32 / 64 output
  channels), cropped by one column, relu, written at a channel offset.  (cout = 64 only merges with
  ADVOC_TC_MERGE_MAX_BN=64; by default it checks the per-class schedule on the same geometry.)"""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from oracle import nets_torch as O
  half = math == 'f16'
  m = N.MATH_F16 if half else N.MATH_AUTO
  q = (lambda t: t.half().float()) if half else _tf32
  adt = torch.float16 if half else torch.float32
  g = torch.Generator().manual_seed(40 + cout)
  B, H, W, Cin = 16, 64, 65, 64
  x = q(torch.relu(torch.randn(B, H, W, Cin, generator=g)))
  k = torch.randn(4, 4, cout, Cin, generator=g) * 0.05
  b = torch.randn(cout, generator=g) * 0.1
  ref = torch.relu(O.deconv_same(x, q(k), b, (2, 2)))[:, :, :-1, :]
  out = torch.full((B, 2 * H, 2 * W - 1, cout + 32), float('nan'), device='cuda', dtype=adt)
  L = nets._Conv('t', 'deconv', nets._desc(B, 2 * H, 2 * W, cout, Cin, 2, 2, 1, 1, H, W, m))
  L.ldx = Cin
  import os
  if not os.environ.get('ADVOC_P2D_FORCE'):      # (tests/test_gpu_forced_paths.py: the patch kernel then takes the layer)
    assert L.kernel_family().startswith('conv_tc')
  xd, kd, bd = x.cuda().to(adt), k.cuda(), b.cuda()
  wp = nets._pack_for_tc(L, kd, Cin)
  ep = nets._epilogue(bd, out, cout + 32, 32, N.ACT_RELU, store_w=2 * W - 1)
  L.run(xd, Cin, wp, ep)
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert _rel(out[..., 32:].float(), ref) < TOL
  assert torch.isnan(out[..., :32].float()).all()


def test_captured_split_k_graph_survives_workspace_growth():
  """A CUDA graph captured around a split-K layer keeps pointing at the library workspace it was captured with;
  a later, larger request (here a 512 x 512-channel filter gradient: 5 splits x 16.8 MB) must not free that buffer
  under the graph (csrc/wgrad_tc.cu: wgrad_workspace retires outgrown buffers instead of freeing them)."""
  import ctypes as C
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  g = torch.Generator().manual_seed(9)
  B, H, W, Cin, Cout = 4, 2, 5, 512, 256
  x = _tf32(torch.randn(B, H, W, Cin, generator=g)).cuda()
  k = (torch.randn(4, 4, Cin, Cout, generator=g) * 0.02).cuda()
  ho, pt, _ = nets.same_pads(H, 4, 2)
  wo, pl, _ = nets.same_pads(W, 4, 2)
  L = nets._Conv('t', 'conv', nets._desc(B, H, W, Cin, Cout, 2, 2, pt, pl, ho, wo, N.MATH_AUTO))
  wp = nets._pack_for_tc(L, k, Cin)
  y = torch.zeros((B, ho, wo, Cout), device='cuda')
  ep = nets._epilogue(None, y, Cout, 0, N.ACT_NONE)
  n0 = N.launch_count()
  L.run(x, Cin, wp, ep)                       # eager warm-up: sizes the workspace
  assert N.launch_count() - n0 == 2          # split-K + finalize
  torch.cuda.synchronize()
  first = y.clone()
  graph = torch.cuda.CUDAGraph()
  with torch.cuda.graph(graph):
    L.run(x, Cin, wp, ep)
  # a request well beyond the 64 MB the workspace starts with
  Bw, Hw = 4, 64
  d = N.ConvDesc(Bw, Hw, Hw, 512, 512, 4, 4, 2, 2, 1, 1, Hw // 2, Hw // 2, N.MATH_AUTO)
  big = torch.randn(Bw, Hw, Hw, 512, generator=g).cuda()
  small = torch.randn(Bw, Hw // 2, Hw // 2, 512, generator=g).cuda()
  dw = torch.zeros(4, 4, 512, 512, device='cuda')
  N.call('advoc_conv2d_wgrad', C.byref(d), nets._ptr(big), 512, nets._ptr(small), 512, nets._ptr(dw), nets._stream())
  torch.cuda.synchronize()
  junk = [torch.randn(16 << 20, device='cuda') for _ in range(4)]     # would land in a freed workspace
  y.zero_()
  graph.replay()
  torch.cuda.synchronize()
  assert N.debug_flags() == 0
  assert torch.equal(y, first)
  del junk
