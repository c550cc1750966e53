"""CPU restatement (PyTorch fp32) of the reference's MelspecGAN stacks and losses.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it.

Follows models/melspecgan/conv2d.py (layers :4-79, generator :82-150, discriminator :153-219) and
models/melspecgan/train.py (:74-111 losses, :117-135 optimisers) with TF's rules:
  * SAME conv k5 s2 on even sizes: pad (1, 2) on both axes
  * SAME conv_transpose k5 s2  == F.conv_transpose2d(stride=2, padding=1)[..., :-1, :-1] (B.2)
  * weights `W` [kh,kw,in,out] (conv) / [kh,kw,out,in] (conv_transpose), own bias `b`
  * tf.layers.batch_normalization defaults: eps 1e-3, batch statistics when training=True
    (biased variance), gamma 1 / beta 0 init.

Parity status: **values unpinned, generator structure pinned** -- the reference has no test, golden vector
or checkpoint for these nets; what it does hold is the exported inference graph
models/melspecgan/infer.meta, whose variable names / layouts, op order, conv_transpose strides, padding and
output sizes, FusedBatchNorm epsilon (1e-3) and decay (0.99; the FUSED op is what hands the
Bessel-corrected batch variance to the moving average), latent size and feats_denorm tail are extracted
into tests/golden/melspecgan_infer_graph.json (scripts/make_melspecgan_graph_fixture.py) and checked
against this file by tests/test_oracle_melspecgan.py.  The discriminator and the losses rest on code
reading.
"""
import torch
import torch.nn.functional as F

from oracle.nets_torch import same_pads, lrelu

BN_EPS = 1e-3   # tf.layers.batch_normalization default (conv2d.py:108,178 pass no arguments)
DIM = 64
Z_DIM = 100     # train.py:15

G_BN = ['G/batch_normalization', 'G/batch_normalization_1', 'G/batch_normalization_2',
        'G/batch_normalization_3']
D_BN = ['D/batch_normalization', 'D/batch_normalization_1', 'D/batch_normalization_2']


def init_params(seed=0, dim=DIM, dtype=torch.float32):
  """N(0, 0.02) weights, zero biases (conv2d.py:7-12,38-48,68-75), BN gamma 1 / beta 0, under the
  variable names of models/melspecgan/infer.meta."""
  g = torch.Generator().manual_seed(seed)
  P = {}

  def w(name, shape, nb):
    P[name + '/W'] = (torch.randn(shape, generator=g, dtype=torch.float64) * 0.02).to(dtype)
    P[name + '/b'] = torch.zeros(nb, dtype=dtype)

  def bn(name, c):
    P[name + '/gamma'] = torch.ones(c, dtype=dtype)
    P[name + '/beta'] = torch.zeros(c, dtype=dtype)

  w('G/z_proj', (Z_DIM, 4 * 5 * dim * 8), 4 * 5 * dim * 8)
  bn(G_BN[0], dim * 8)
  chans = [dim * 8, dim * 4, dim * 2, dim, 1]
  for i in range(4):
    w('G/upconv_%d' % (i + 1), (5, 5, chans[i + 1], chans[i]), chans[i + 1])
    if i < 3:
      bn(G_BN[i + 1], chans[i + 1])
  dch = [1, dim, dim * 2, dim * 4, dim * 8]
  for i in range(4):
    w('D/conv_%d' % i, (5, 5, dch[i], dch[i + 1]), dch[i + 1])
    if i > 0:
      bn(D_BN[i - 1], dch[i + 1])
  w('D/out', (4 * 5 * dim * 8, 1), 1)
  return P


BN_MOMENTUM = 0.99   # tf.layers.batch_normalization default


def init_moving(dim=DIM, dtype=torch.float32):
  """moving_mean 0 / moving_variance 1 of the generator's four batch norms (infer.meta variable names)."""
  M = {}
  for name, c in zip(G_BN, [dim * 8, dim * 4, dim * 2, dim]):
    M[name + '/moving_mean'] = torch.zeros(c, dtype=dtype)
    M[name + '/moving_variance'] = torch.ones(c, dtype=dtype)
  return M


def batchnorm(x, gamma, beta, moving=None, name=None, training=True):
  """training=True: batch statistics over (N, H, W) per channel, biased variance; when `moving` is given
  its entries are updated in place the way the layer's UPDATE_OPS do (conv2d.py:143-148):
  `moving -= (moving - batch) * (1 - 0.99)`, the batch variance with Bessel's correction n/(n-1) (what
  TF's fused batch-norm op hands to the moving average for 4-D inputs).  training=False: normalise
  with the moving statistics (models/melspecgan/infer.py:17)."""
  if not training:
    m, v = moving[name + '/moving_mean'], moving[name + '/moving_variance']
    return (x - m) * torch.rsqrt(v + BN_EPS) * gamma + beta
  m = x.mean(dim=(0, 1, 2), keepdim=True)
  v = ((x - m) ** 2).mean(dim=(0, 1, 2), keepdim=True)
  if moving is not None:
    n = x.numel() // x.shape[-1]
    with torch.no_grad():
      mm, mv = moving[name + '/moving_mean'], moving[name + '/moving_variance']
      mm -= (mm - m.reshape(-1)) * (1. - BN_MOMENTUM)
      mv -= (mv - v.reshape(-1) * (n / (n - 1.))) * (1. - BN_MOMENTUM)
  return (x - m) * torch.rsqrt(v + BN_EPS) * gamma + beta


def conv5(x, W, b):
  xt = x.permute(0, 3, 1, 2)
  pt, pb = same_pads(xt.shape[2], 5, 2)
  pl, pr = same_pads(xt.shape[3], 5, 2)
  y = F.conv2d(F.pad(xt, (pl, pr, pt, pb)), W.permute(3, 2, 0, 1), b, stride=2)
  return y.permute(0, 2, 3, 1)


def deconv5(x, W, b):
  xt = x.permute(0, 3, 1, 2)
  y = F.conv_transpose2d(xt, W.permute(3, 2, 0, 1), b, stride=2, padding=1)[:, :, :-1, :-1]
  return y.permute(0, 2, 3, 1)


def generator(P, z, dim=DIM, return_layers=False, moving=None, training=True):
  """z [b, 100] -> [b, 64, 80, 1] in (-1, 1)  (conv2d.py:95-150).  `moving`: see batchnorm."""
  x = z @ P['G/z_proj/W'] + P['G/z_proj/b']
  x = x.reshape(-1, 4, 5, dim * 8)
  x = torch.relu(batchnorm(x, P[G_BN[0] + '/gamma'], P[G_BN[0] + '/beta'], moving, G_BN[0], training))
  layers = [x]
  for i in range(1, 5):
    x = deconv5(x, P['G/upconv_%d/W' % i], P['G/upconv_%d/b' % i])
    if i < 4:
      x = torch.relu(batchnorm(x, P[G_BN[i] + '/gamma'], P[G_BN[i] + '/beta'], moving, G_BN[i], training))
    else:
      x = torch.tanh(x)
    layers.append(x)
  return (x, layers) if return_layers else x


def discriminator(P, x, return_layers=False, dim=None):
  """x [b, 64, 80, 1] -> logits [b]  (conv2d.py:166-219, training=True)."""
  layers = []
  for i in range(4):
    x = conv5(x, P['D/conv_%d/W' % i], P['D/conv_%d/b' % i])
    if i > 0:
      x = batchnorm(x, P[D_BN[i - 1] + '/gamma'], P[D_BN[i - 1] + '/beta'])
    x = lrelu(x)
    layers.append(x)
  x = x.reshape(x.shape[0], -1) @ P['D/out/W'] + P['D/out/b']
  out = x[:, 0]
  return (out, layers) if return_layers else out


def losses(P, z, x, train_loss='wgangp', alpha=None):
  """G_loss, D_loss of train.py:74-111.  `alpha` [b,1,1,1] = the interpolation draw of :100."""
  return losses_dim(P, z, x, P['G/upconv_4/W'].shape[3], train_loss, alpha)


def losses_dim(P, z, x, dim, train_loss='wgangp', alpha=None):
  G_z = generator(P, z, dim)
  D_x = discriminator(P, x)
  D_G_z = discriminator(P, G_z)
  if train_loss == 'dcgan':
    ones, zeros = torch.ones_like(D_x), torch.zeros_like(D_x)
    G_loss = F.binary_cross_entropy_with_logits(D_G_z, ones)
    D_loss = (F.binary_cross_entropy_with_logits(D_G_z, zeros) +
              F.binary_cross_entropy_with_logits(D_x, ones)) / 2.
    return dict(G_loss=G_loss, D_loss=D_loss, G_z=G_z, D_x=D_x, D_G_z=D_G_z)
  G_loss = -D_G_z.mean()
  D_loss = D_G_z.mean() - D_x.mean()
  interp = x + alpha * (G_z - x)
  if not interp.requires_grad:
    interp = interp.detach().requires_grad_(True)   # D-only gradients are wanted: G_z is a constant here
  D_i = discriminator(P, interp)
  grad = torch.autograd.grad(D_i.sum(), interp, create_graph=True)[0]
  slopes = torch.sqrt((grad ** 2).sum(dim=(1, 2, 3)))
  gp = ((slopes - 1.) ** 2).mean()
  return dict(G_loss=G_loss, D_loss=D_loss + 10. * gp, gp=gp, G_z=G_z, D_x=D_x, D_G_z=D_G_z)


def g_names(P):
  return [n for n in P if n.startswith('G/')]


def d_names(P):
  return [n for n in P if n.startswith('D/')]
