"""CPU restatement (PyTorch fp32) of the reference's AdVoc conv stacks.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product path (`advoc_b200/`) never does.

The reference builds these nets out of stock TensorFlow 1.13 ops
(tensorflow-gpu<=1.13.1, setup.py:17), which is not installable in this image,
so this file restates the graph with torch CPU ops using TF's padding rules
(verified with autograd identities):
  * SAME conv k4 s2      : pad (lo, hi) = (1, 1) for even input, (1, 2) for odd input
  * SAME conv_transpose  : == F.conv_transpose2d(stride=2, padding=1), kernel HWOI, no flip
  * explicit pad 1 + VALID for the discriminator

Parity status: **parity unpinned** -- the reference holds no test, golden vector
or checkpoint for the conv stacks (tests/ only cover advoc.audioio/advoc.spectral),
so the fidelity of this restatement to TF rests on code reading and appendix B.

Tensors are NHWC [b, time, freq, ch] like the reference; parameters are a flat
dict keyed by the TF variable names, kernels in TF layouts (HWIO conv, HWOI deconv).
"""
import math

import torch
import torch.nn.functional as F

EPS = 1e-12  # models/advoc/advoc_model.py:8


# ---------------------------------------------------------------------------
# model hyper-parameters (models/advoc/advoc_model.py:11-22, advoc_model_small.py:14-23)
# ---------------------------------------------------------------------------
class Spec(object):
  def __init__(self, ngf, ndf, num_enc_layers, dropout_decoders, subseq_len=256):
    self.ngf, self.ndf = ngf, ndf
    self.num_enc_layers = num_enc_layers      # encoders after encoder_1
    self.dropout_decoders = dropout_decoders  # decoder indices with dropout(keep .5)
    self.subseq_len = subseq_len


REGULAR = Spec(64, 64, 7, (8, 7, 6))
SMALL = Spec(32, 32, 4, (5, 4))


def encoder_channels(spec):
  mult = [1, 2, 4, 8, 8, 8, 8, 8][:spec.num_enc_layers + 1]
  return [spec.ngf * m for m in mult]


def decoder_channels(spec):
  """Output channels of decoder_k for k = n_enc .. 2 (decoder_1 -> 1 channel)."""
  full = {8: 8, 7: 8, 6: 8, 5: 8, 4: 4, 3: 2, 2: 1}
  n_enc = spec.num_enc_layers + 1
  return {k: spec.ngf * full[k] for k in range(n_enc, 1, -1)}


def init_params(spec, seed=0, dtype=torch.float32):
  """N(0, 0.02) kernels, zero biases (advoc_model.py:32,36,55; tf.layers default bias)."""
  g = torch.Generator().manual_seed(seed)
  P = {}

  def kern(name, shape):
    P[name + '/kernel'] = (torch.randn(shape, generator=g, dtype=torch.float64) * 0.02).to(dtype)
    P[name + '/bias'] = torch.zeros(shape[3] if 'conv2d_transpose' not in name else shape[2],
                                    dtype=dtype)

  enc = encoder_channels(spec)
  cin = 1
  for i, c in enumerate(enc):
    kern('generator/encoder_%d/conv2d' % (i + 1), (4, 4, cin, c))
    cin = c
  dec = decoder_channels(spec)
  n_enc = len(enc)
  prev = enc[-1]
  for k in range(n_enc, 1, -1):
    cin = prev if k == n_enc else prev + enc[k - 1]
    kern('generator/decoder_%d/conv2d_transpose' % k, (4, 4, dec[k], cin))
    prev = dec[k]
  kern('generator/decoder_1/conv2d_transpose', (4, 4, 1, prev + enc[0]))
  cin = 2
  for i, c in enumerate([spec.ndf, spec.ndf * 2, spec.ndf * 4, spec.ndf * 8, 1]):
    kern('discriminator/layer_%d/conv2d' % (i + 1), (4, 4, cin, c))
    cin = c
  return P


# ---------------------------------------------------------------------------
# TF-exact layer primitives
# ---------------------------------------------------------------------------
def same_pads(n, k, s):
  out = -(-n // s)
  total = max((out - 1) * s + k - n, 0)
  return total // 2, total - total // 2


def conv_same(x, kernel, bias, strides=(2, 2)):
  """tf.layers.conv2d(k=4, padding='same') on NHWC; kernel HWIO (advoc_model.py:46-51)."""
  xt = x.permute(0, 3, 1, 2)
  pt, pb = same_pads(xt.shape[2], kernel.shape[0], strides[0])
  pl, pr = same_pads(xt.shape[3], kernel.shape[1], strides[1])
  xt = F.pad(xt, (pl, pr, pt, pb))
  # (.contiguous(): torch's float64 CPU conv backward insists on a contiguous weight gradient)
  y = F.conv2d(xt, kernel.permute(3, 2, 0, 1).contiguous(), bias, stride=strides)
  return y.permute(0, 2, 3, 1)


def deconv_same(x, kernel, bias, strides=(2, 2)):
  """tf.layers.conv2d_transpose(k=4, s=2, 'same'); kernel HWOI (advoc_model.py:65-69)."""
  xt = x.permute(0, 3, 1, 2)
  y = F.conv_transpose2d(xt, kernel.permute(3, 2, 0, 1).contiguous(), bias, stride=strides, padding=1)
  if strides[0] == 1:
    y = y[:, :, :-1, :]
  return y.permute(0, 2, 3, 1)


def discrim_conv(x, kernel, bias, stride):
  """tf.pad 1 + conv2d VALID (advoc_model.py:25-32)."""
  xt = F.pad(x.permute(0, 3, 1, 2), (1, 1, 1, 1))
  y = F.conv2d(xt, kernel.permute(3, 2, 0, 1).contiguous(), bias, stride=stride)
  return y.permute(0, 2, 3, 1)


def lrelu(x, alpha=0.2):
  return torch.maximum(alpha * x, x)


# ---------------------------------------------------------------------------
# generator / discriminator  (advoc_model.py:75-166, :168-204)
# ---------------------------------------------------------------------------
def generator(P, x, spec, dropout_masks=None, return_layers=False):
  """x [b, T, 513, 1] -> [b, T, 513, 1].

  dropout_masks: None -> dropout disabled (parity mode); else dict
  {decoder_index: 0/1 mask tensor of that decoder's output shape}; kept values are
  scaled by 1/keep_prob = 2 (tf.nn.dropout, advoc_model.py:144-149).
  """
  n_enc = spec.num_enc_layers + 1
  layers = []
  n_time = spec.subseq_len
  out = conv_same(x, P['generator/encoder_1/conv2d/kernel'], P['generator/encoder_1/conv2d/bias'])
  n_time //= 2
  layers.append(out)
  n_stride1 = 0
  for i in range(2, n_enc + 1):
    name = 'generator/encoder_%d/conv2d' % i
    rect = lrelu(layers[-1], 0.2)
    if n_time > 1:
      out = conv_same(rect, P[name + '/kernel'], P[name + '/bias'], (2, 2))
      n_time //= 2
    else:
      n_stride1 += 1
      out = conv_same(rect, P[name + '/kernel'], P[name + '/bias'], (1, 2))
    layers.append(out)
  for j, k in enumerate(range(n_enc, 1, -1)):
    name = 'generator/decoder_%d/conv2d_transpose' % k
    skip = k - 1  # index into layers of encoder_k
    if j == 0:
      inp = layers[-1]
    else:
      inp = torch.cat([layers[-1][:, :, :-1, :], layers[skip]], dim=3)
    rect = torch.relu(inp)
    strides = (1, 2) if j < n_stride1 else (2, 2)
    out = deconv_same(rect, P[name + '/kernel'], P[name + '/bias'], strides)
    if k in spec.dropout_decoders and dropout_masks is not None:
      out = out * dropout_masks[k] * 2.0
    layers.append(out)
  inp = torch.cat([layers[-1][:, :, :-1, :], layers[0]], dim=3)
  rect = torch.relu(inp)
  name = 'generator/decoder_1/conv2d_transpose'
  out = deconv_same(rect, P[name + '/kernel'], P[name + '/bias'])[:, :, :-1, :]
  layers.append(out)
  return (out, layers) if return_layers else out


def discriminator(P, inputs, targets, return_layers=False):
  """[b,T,513,1] x2 -> sigmoid patch map [b,30,62,1] (advoc_model.py:168-204)."""
  x = torch.cat([inputs, targets], dim=3)
  layers = []
  strides = [2, 2, 2, 1, 1]
  for i in range(5):
    name = 'discriminator/layer_%d/conv2d' % (i + 1)
    x = discrim_conv(x, P[name + '/kernel'], P[name + '/bias'], strides[i])
    x = lrelu(x, 0.2) if i < 4 else torch.sigmoid(x)
    layers.append(x)
  return (x, layers) if return_layers else x


# ---------------------------------------------------------------------------
# losses (advoc_model.py:238-245)
# ---------------------------------------------------------------------------
def losses(P, x, target, spec, gan_weight=1.0, l1_weight=10.0, dropout_masks=None):
  gen = generator(P, x, spec, dropout_masks)
  p_real = discriminator(P, x, target)
  p_fake = discriminator(P, x, gen)
  d_loss = torch.mean(-(torch.log(p_real + EPS) + torch.log(1 - p_fake + EPS)))
  g_gan = torch.mean(-torch.log(p_fake + EPS))
  g_l1 = torch.mean(torch.abs(target - gen))
  g_loss = g_gan * gan_weight + g_l1 * l1_weight if gan_weight > 0 else g_l1 * l1_weight
  return dict(gen=gen, d_loss=d_loss, g_gan=g_gan, g_l1=g_l1, g_loss=g_loss)


# ---------------------------------------------------------------------------
# TF1 Adam (tf.train.AdamOptimizer(2e-4, 0.5); SURVEY appendix B rule 5)
# ---------------------------------------------------------------------------
class TFAdam(object):
  def __init__(self, names, P, lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8):
    self.names = list(names)
    self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
    self.t = 0
    self.m = {n: torch.zeros_like(P[n]) for n in self.names}
    self.v = {n: torch.zeros_like(P[n]) for n in self.names}

  def step(self, P, grads):
    self.t += 1
    lr_t = self.lr * math.sqrt(1 - self.b2 ** self.t) / (1 - self.b1 ** self.t)
    for n in self.names:
      g = grads[n]
      self.m[n].mul_(self.b1).add_(g, alpha=1 - self.b1)
      self.v[n].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
      P[n] = P[n] - lr_t * self.m[n] / (self.v[n].sqrt() + self.eps)


def g_names(P):
  return [n for n in P if n.startswith('generator')]


def d_names(P):
  return [n for n in P if n.startswith('discriminator')]


def grads_of(loss, P, names):
  leaves = [P[n] for n in names]
  gs = torch.autograd.grad(loss, leaves, allow_unused=False)
  return dict(zip(names, gs))


def train_step(P, opt_d, opt_g, batch_d, batch_g, spec, masks_d=None, masks_g=None,
               gan_weight=1.0, l1_weight=10.0):
  """One `train_loop` (advoc_model.py:285-289): D update on batch_d, then G update on
  batch_g through the already-updated D.  batch = (x, target)."""
  out = {}
  if gan_weight > 0:
    Pd = {n: (t.detach().requires_grad_(n.startswith('discriminator'))) for n, t in P.items()}
    l = losses(Pd, batch_d[0], batch_d[1], spec, gan_weight, l1_weight, masks_d)
    gd = grads_of(l['d_loss'], Pd, d_names(P))
    opt_d.step(P, gd)
    out['d_loss'] = float(l['d_loss'])
    out['d_grads'] = gd
  Pg = {n: (t.detach().requires_grad_(n.startswith('generator'))) for n, t in P.items()}
  l = losses(Pg, batch_g[0], batch_g[1], spec, gan_weight, l1_weight, masks_g)
  gg = grads_of(l['g_loss'], Pg, g_names(P))
  opt_g.step(P, gg)
  out.update(g_loss=float(l['g_loss']), g_gan=float(l['g_gan']), g_l1=float(l['g_l1']),
             g_grads=gg)
  return out


def count_params(P, prefix):
  return sum(t.numel() for n, t in P.items() if n.startswith(prefix))
