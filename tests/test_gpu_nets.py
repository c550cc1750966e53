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
