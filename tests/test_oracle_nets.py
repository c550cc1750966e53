"""Checks of the PyTorch-CPU restatement of the conv stacks (oracle/nets_torch.py) against the
layer tables derived from the reference.  The reference holds no
golden for these nets ("parity unpinned"), so what can be pinned is: shapes, parameter counts,
the TF SAME-padding rule and the conv/conv_transpose adjoint identity.  CPU only.
"""
import torch
import torch.nn.functional as F

from oracle import nets_torch as O


def test_param_counts():
  # SURVEY appendix A.1/A.2 (derived from advoc_model.py:91-158,181-204)
  P = O.init_params(O.REGULAR)
  assert O.count_params(P, 'generator') == 54403457
  assert O.count_params(P, 'discriminator') == 2763713
  P = O.init_params(O.SMALL)
  assert O.count_params(P, 'generator') == 4164289
  assert O.count_params(P, 'discriminator') == 693729


def test_same_pads_rule():
  assert O.same_pads(256, 4, 2) == (1, 1)
  assert O.same_pads(513, 4, 2) == (1, 2)
  assert O.same_pads(257, 4, 2) == (1, 2)
  assert O.same_pads(80, 5, 2) == (1, 2)
  assert O.same_pads(1, 4, 1) == (1, 2)


def test_small_generator_shapes():
  P = O.init_params(O.SMALL)
  x = torch.randn(1, 256, 513, 1)
  out, layers = O.generator(P, x, O.SMALL, return_layers=True)
  shapes = [tuple(l.shape[1:]) for l in layers]
  assert shapes == [(128, 257, 32), (64, 129, 64), (32, 65, 128), (16, 33, 256), (8, 17, 256),
                    (16, 34, 256), (32, 66, 128), (64, 130, 64), (128, 258, 32), (256, 513, 1)]
  assert out.shape == (1, 256, 513, 1)


def test_discriminator_shapes():
  P = O.init_params(O.SMALL)
  x = torch.randn(1, 256, 513, 1)
  out, layers = O.discriminator(P, x, x, return_layers=True)
  assert [tuple(l.shape[1:]) for l in layers] == [(128, 256, 32), (64, 128, 64), (32, 64, 128),
                                                 (31, 63, 256), (30, 62, 1)]
  assert float(out.min()) > 0 and float(out.max()) < 1


def test_deconv_is_adjoint_of_same_conv():
  # SURVEY appendix B rule 2: SAME conv_transpose == input-gradient of the SAME conv
  torch.manual_seed(0)
  k = torch.randn(4, 4, 3, 5, dtype=torch.float64)      # HWOI: out 3, in 5
  x = torch.randn(2, 6, 9, 5, dtype=torch.float64)      # small side
  y = O.deconv_same(x, k, None)                         # [2,12,18,3]
  assert y.shape == (2, 12, 18, 3)
  big = torch.randn(2, 12, 18, 3, dtype=torch.float64, requires_grad=True)
  # forward SAME conv of the big side with the same kernel read as HWIO (in 3, out 5)
  z = O.conv_same(big, k, None)
  assert z.shape == x.shape
  (g,) = torch.autograd.grad((z * x).sum(), big)
  assert torch.allclose(g, y, atol=1e-10)


def test_short_subseq_uses_stride1_layers():
  spec = O.Spec(8, 8, 7, (8, 7, 6), subseq_len=64)
  P = O.init_params(spec)
  x = torch.randn(1, 64, 513, 1)
  out, layers = O.generator(P, x, spec, return_layers=True)
  assert out.shape == (1, 64, 513, 1)
  assert tuple(layers[6].shape[1:3]) == (1, 5) and tuple(layers[7].shape[1:3]) == (1, 3)


def test_tf_adam_matches_formula():
  P = {'generator/x': torch.tensor([1.0, -2.0])}
  opt = O.TFAdam(['generator/x'], P, lr=0.1, beta1=0.5, beta2=0.999, eps=1e-8)
  g = torch.tensor([0.5, -0.25])
  opt.step(P, {'generator/x': g})
  m = 0.5 * g
  v = 0.001 * g * g
  lr_t = 0.1 * (1 - 0.999) ** 0.5 / (1 - 0.5)
  exp = torch.tensor([1.0, -2.0]) - lr_t * m / (v.sqrt() + 1e-8)
  assert torch.allclose(P['generator/x'], exp)
