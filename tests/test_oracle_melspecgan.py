"""CPU checks of the MelspecGAN oracle (oracle/melspecgan_torch.py): layer shapes and parameter
counts of SURVEY.md appendix A.3, the TF SAME-padding identities of appendix B for k5 s2, batch
norm against torch's own, and that both reference losses (train.py:76-111) differentiate."""
import torch
import torch.nn.functional as F

from oracle import melspecgan_torch as M


def test_shapes_and_param_counts():
  P = M.init_params(seed=0)
  g = torch.Generator().manual_seed(1)
  z = torch.randn(3, M.Z_DIM, generator=g)
  G_z, layers = M.generator(P, z, return_layers=True)
  assert [tuple(l.shape[1:]) for l in layers] == [(4, 5, 512), (8, 10, 256), (16, 20, 128), (32, 40, 64),
                                                  (64, 80, 1)]
  assert float(G_z.abs().max()) < 1.0
  out, dl = M.discriminator(P, G_z, return_layers=True)
  assert out.shape == (3,)
  assert [tuple(l.shape[1:]) for l in dl] == [(32, 40, 64), (16, 20, 128), (8, 10, 256), (4, 5, 512)]
  n_g = sum(P[n].numel() for n in M.g_names(P) if n.endswith('/W') or n.endswith('/b'))
  n_d = sum(P[n].numel() for n in M.d_names(P) if n.endswith('/W') or n.endswith('/b'))
  assert n_g == 5337089 and n_d == 4313601


def test_k5_same_padding_identities():
  """conv SAME k5 s2 pads (1, 2); conv_transpose SAME is its adjoint (== the input gradient)."""
  g = torch.Generator().manual_seed(2)
  x = torch.randn(2, 8, 10, 3, generator=g, requires_grad=True)
  W = torch.randn(5, 5, 3, 4, generator=g)
  y = M.conv5(x, W, None)
  assert y.shape == (2, 4, 5, 4)
  dy = torch.randn(y.shape, generator=g)
  (gx,) = torch.autograd.grad(y, x, dy)
  # deconv5 takes the conv_transpose layout [kh,kw,out,in] = the conv's HWIO with roles swapped
  assert torch.allclose(M.deconv5(dy, W, None), gx, atol=1e-5)


def test_batchnorm_matches_torch():
  g = torch.Generator().manual_seed(3)
  x = torch.randn(4, 6, 5, 8, generator=g) * 3 + 1
  gamma, beta = torch.randn(8, generator=g), torch.randn(8, generator=g)
  ref = F.batch_norm(x.permute(0, 3, 1, 2), None, None, gamma, beta, training=True, eps=M.BN_EPS).permute(0, 2, 3, 1)
  assert torch.allclose(M.batchnorm(x, gamma, beta), ref, atol=1e-5)


def test_losses_differentiate():
  P = {n: t.clone().requires_grad_(True) for n, t in M.init_params(seed=0, dim=8).items()}
  g = torch.Generator().manual_seed(4)
  z = torch.randn(4, M.Z_DIM, generator=g)
  x = torch.rand(4, 64, 80, 1, generator=g) * 2 - 1
  for kind in ('dcgan', 'wgangp'):
    l = M.losses_dim(P, z, x, 8, kind, alpha=torch.rand(4, 1, 1, 1, generator=g))
    gd = torch.autograd.grad(l['D_loss'], [P[n] for n in M.d_names(P)], retain_graph=True, allow_unused=True)
    gg = torch.autograd.grad(l['G_loss'], [P[n] for n in M.g_names(P)], allow_unused=True)
    assert all(t is not None and torch.isfinite(t).all() for t in gd)
    assert all(t is not None and torch.isfinite(t).all() for t in gg)


def test_moving_average_update_rule():
  """One training-mode evaluation from the TF initial values (0, 1): moving -= (moving - batch) * 0.01 with
  the Bessel-corrected batch variance; training=False then reads them and a larger batch does not matter."""
  import torch
  from oracle import melspecgan_torch as M
  P = M.init_params(seed=1, dim=8)
  g = torch.Generator().manual_seed(2)
  z = torch.randn(5, M.Z_DIM, generator=g)
  moving = M.init_moving(8)
  M.generator(P, z, 8, moving=moving)
  x0 = (z @ P['G/z_proj/W'] + P['G/z_proj/b']).reshape(-1, 4, 5, 64).reshape(-1, 64)
  assert torch.allclose(moving['G/batch_normalization/moving_mean'], 0.01 * x0.mean(0), atol=1e-7)
  assert torch.allclose(moving['G/batch_normalization/moving_variance'], 0.99 + 0.01 * x0.var(0, unbiased=True),
                        atol=1e-7)
  one = M.generator(P, z[:1], 8, moving=moving, training=False)
  many = M.generator(P, z, 8, moving=moving, training=False)
  assert torch.allclose(one, many[:1], atol=1e-6)      # no batch coupling in the inference graph
