"""MelspecGAN on B200: DCGAN-style generator / discriminator over 64 x 80 mel patches, forward,
backward and TF1-Adam, data-parallel like the AdVoc train step.

reference: models/melspecgan/conv2d.py (`dense_layer` :4-14, `conv2d_transpose_layer` :17-52,
`conv2d_layer` :55-79, `MelspecGANGenerator.__call__` :95-150, `MelspecGANDiscriminator.__call__`
:166-219) and models/melspecgan/train.py (:44-57 z / G, :64-72 D on real and fake, :74-111 losses,
:117-153 optimisers and the D:G schedule).

Every op is a launch into libadvoc_b200.so on persistent NHWC buffers: the 5x5 stride-2
convolutions through the same conv entry points as AdVoc (tcgen05 TF32 where the channel counts
allow it), dense layers / batch normalisation / tanh' / logit losses through csrc/melspecgan.cu.
Batch normalisation uses batch statistics (`training=True`, the only mode the reference trains
in) over each discriminator call's own batch, as the reference's three separate `D(...)` graphs do.

Parameters are a dict keyed by the reference's variable names (`G/z_proj/{W,b}`,
`G/upconv_{1..4}/{W,b}`, `G/batch_normalization{,_1,_2,_3}/{gamma,beta}`, `D/conv_{0..3}/{W,b}`,
`D/batch_normalization{,_1,_2}/{gamma,beta}`, `D/out/{W,b}`), weights in the TF layouts.

All three losses of the reference are built: 'dcgan' (train.py:76-94), 'wgangp' (:96-111, the
default: critic loss + 10 x gradient penalty on interpolates, 5 D steps per G step) and, for
ablation, 'wgan' (the critic terms without the penalty).  The penalty's parameter gradient is the
reference's `tf.gradients` inside the loss, i.e. a double backward through conv + batch-norm +
leaky-ReLU; it is laid out explicitly in `_gp_backward`: an "up" sweep (the adjoint of the first
backward pass: forward convs of the adjoint, the batch-norm second-order terms of
`advoc_bn_gp`) and a "down" sweep (an ordinary backward pass seeded with those terms).
"""
import ctypes as C

import torch

from advoc_b200 import _native as N
from advoc_b200 import nets
from advoc_b200.nets import _epilogue, _ptr, _stream
from advoc_b200.train import FlatParams

BN_EPS = 1e-3   # tf.layers.batch_normalization default
Z_DIM = 100     # train.py:15
G_BN = ['G/batch_normalization', 'G/batch_normalization_1', 'G/batch_normalization_2',
        'G/batch_normalization_3']
D_BN = ['D/batch_normalization', 'D/batch_normalization_1', 'D/batch_normalization_2']


def init_params(seed=0, dim=64, device='cuda'):
  """N(0, 0.02) weights, zero biases, gamma 1 / beta 0 (conv2d.py:7-12,38-48,68-75)."""
  g = torch.Generator(device='cpu').manual_seed(seed)
  P = {}

  def w(name, shape, nb):
    P[name + '/W'] = (torch.randn(shape, generator=g) * 0.02).to(device)
    P[name + '/b'] = torch.zeros(nb, device=device)

  def bn(name, c):
    P[name + '/gamma'] = torch.ones(c, device=device)
    P[name + '/beta'] = torch.zeros(c, device=device)

  w('G/z_proj', (Z_DIM, 4 * 5 * dim * 8), 4 * 5 * dim * 8)
  bn(G_BN[0], dim * 8)
  ch = [dim * 8, dim * 4, dim * 2, dim, 1]
  for i in range(4):
    w('G/upconv_%d' % (i + 1), (5, 5, ch[i + 1], ch[i]), ch[i + 1])
    if i < 3:
      bn(G_BN[i + 1], ch[i + 1])
  dch = [1, dim, dim * 2, dim * 4, dim * 8]
  for i in range(4):
    w('D/conv_%d' % i, (5, 5, dch[i], dch[i + 1]), dch[i + 1])
    if i > 0:
      bn(D_BN[i - 1], dch[i + 1])
  w('D/out', (4 * 5 * dim * 8, 1), 1)
  return P


def _desc5(n, h, w, cin, cout, math):
  """5x5 stride-2 SAME conv on even sizes: TF pads (1, 2) on both axes."""
  return N.ConvDesc(n, h, w, cin, cout, 5, 5, 2, 2, 1, 1, h // 2, w // 2, math)


class _DBufs(object):
  """Activations of one discriminator call (real or fake batch)."""

  def __init__(self, B, dch, dev):
    f32 = dict(dtype=torch.float32, device=dev)
    h, w = 64, 80
    self.X, self.Y, self.dX, self.dY, self.stats, self.red = {}, {}, {}, {}, {}, {}
    for i in range(4):
      h, w = h // 2, w // 2
      shape = (B, h, w, dch[i + 1])
      self.Y[i] = torch.empty(shape, **f32)
      self.dY[i] = torch.empty(shape, **f32)
      self.dX[i] = torch.empty(shape, **f32)
      if i > 0:
        self.X[i] = torch.empty(shape, **f32)
        self.stats[i] = torch.zeros(2 * dch[i + 1], **f32)
        self.red[i] = torch.zeros(2 * dch[i + 1], **f32)
    self.logits = torch.empty(B, **f32)
    self.dlogits = torch.empty(B, **f32)

  def alloc_gp(self):
    """Extra buffers of the gradient-penalty double backward (interpolates pass only)."""
    self.v, self.vy, self.xb, self.tmp, self.sums = {}, {}, {}, {}, {}
    for i in range(4):
      self.vy[i] = torch.empty_like(self.Y[i])      # adjoint of the first backward's dY_i (masked)
      self.xb[i] = torch.empty_like(self.Y[i])      # adjoint of the forward X_i
      if i > 0:
        self.v[i] = torch.empty_like(self.Y[i])     # adjoint of the first backward's dX_i
        self.tmp[i] = torch.empty_like(self.Y[i])
        self.sums[i] = torch.zeros(5 * self.Y[i].shape[-1], dtype=torch.float32, device=self.Y[i].device)


class MelspecGAN(object):
  """One replica of the MelspecGAN train step for a fixed per-GPU batch."""

  def __init__(self, params, batch, dim=64, train_loss='dcgan', math=N.MATH_AUTO, process_group=None,
               world_size=1, use_graphs=True, base_seed=1234, rank=0):
    """`use_graphs`: from the second call on, the gradient computation and the optimiser update of
    `d_step` / `g_step` replay from captured CUDA graphs (the ~300 / ~250 launches of one step are
    otherwise bound by the host's launch rate); the gradient all-reduce between the two stays eager."""
    if train_loss not in ('dcgan', 'wgan', 'wgangp'):
      raise ValueError()
    self.use_graphs, self._graphs, self._graph_launches = use_graphs, {}, {}
    self.capture_launches = self.replayed_launches = 0    # launch accounting for bench.py
    self.loss_kind, self.B, self.dim, self.math = train_loss, batch, dim, math
    self.pg, self.world = process_group, world_size
    # optimiser settings of train.py:117-135
    if train_loss == 'dcgan':
      self.lr, self.b1, self.b2 = 2e-4, 0.5, 0.999
    else:
      self.lr, self.b1, self.b2 = 1e-4, 0.5, 0.9
    self.eps = 1e-8
    self.flat = FlatParams(params, gen_prefix='G/', dis_prefix='D/')
    P = self.P = self.flat.P
    dev = self.flat.p.device
    f32 = dict(dtype=torch.float32, device=dev)
    B = batch
    self.rnd = 0 if math == N.MATH_FP32 else 1
    # ---- generator buffers: X[i] pre-BN, Y[i] = relu(BN(X[i])); level 0 is the projected z
    self.gch = [dim * 8, dim * 4, dim * 2, dim, 1]
    self.z = torch.zeros((B, Z_DIM), **f32)
    self.ones = torch.ones((1, B), **f32)
    self.gX, self.gY, self.gdX, self.gdY, self.gstats, self.gred = {}, {}, {}, {}, {}, {}
    h, w = 4, 5
    for i in range(4):
      shape = (B, h, w, self.gch[i])
      self.gX[i], self.gY[i] = torch.empty(shape, **f32), torch.empty(shape, **f32)
      self.gdX[i], self.gdY[i] = torch.empty(shape, **f32), torch.empty(shape, **f32)
      self.gstats[i] = torch.zeros(2 * self.gch[i], **f32)
      self.gred[i] = torch.zeros(2 * self.gch[i], **f32)
      h, w = h * 2, w * 2
    # moving averages of the generator's batch norms (UPDATE_OPS of conv2d.py:143-148, read by the
    # training=False graph of models/melspecgan/infer.py:17); TF initialises them to 0 / 1
    self.bn_momentum = 0.99
    self.g_moving = {i: (torch.zeros(self.gch[i], **f32), torch.ones(self.gch[i], **f32)) for i in range(4)}
    self.G_z = torch.empty((B, 64, 80, 1), **f32)
    self.dG_z = torch.empty((B, 64, 80, 1), **f32)
    self.dX4 = torch.empty((B, 64, 80, 1), **f32)
    # upconv_i (i = 1..4): transposed conv == input gradient of a conv from the big side to the small
    self.up, self.up_b = {}, {}
    h, w = 4, 5
    for i in range(1, 5):
      d = _desc5(B, 2 * h, 2 * w, self.gch[i], self.gch[i - 1], math)
      self.up[i] = nets._Conv('G/upconv_%d' % i, 'deconv', d)     # forward
      self.up_b[i] = nets._Conv('G/upconv_%d' % i, 'conv', d)     # its input gradient
      h, w = 2 * h, 2 * w
    # ---- discriminator
    self.dch = [1, dim, dim * 2, dim * 4, dim * 8]
    self.conv, self.conv_t = {}, {}
    h, w = 64, 80
    for i in range(4):
      d = _desc5(B, h, w, self.dch[i], self.dch[i + 1], math)
      self.conv[i] = nets._Conv('D/conv_%d' % i, 'conv', d)
      self.conv_t[i] = nets._Conv('D/conv_%d' % i, 'deconv', d)
      h, w = h // 2, w // 2
    self.real, self.fake = _DBufs(B, self.dch, dev), _DBufs(B, self.dch, dev)
    self.gp_lambda = 10.0                      # LAMBDA, train.py:104
    if train_loss == 'wgangp':
      self.interp = _DBufs(B, self.dch, dev)
      self.interp.alloc_gp()
      self.x_hat = torch.zeros((B, 64, 80, 1), **f32)
      self.g_hat = torch.zeros((B, 64, 80, 1), **f32)
      self.u_hat = torch.zeros((B, 64, 80, 1), **f32)
      self._alpha_gen = torch.Generator(device=dev)
      self._alpha_gen.manual_seed(base_seed + rank)   # per-replica interpolation alphas under data parallelism
    self.x_real = torch.zeros((B, 64, 80, 1), **f32)
    self.losses = torch.zeros(2, **f32)     # D_loss, G_loss
    self.alpha_buf = torch.zeros((B, 1, 1, 1), **f32)
    self.lr_t = torch.zeros(2, **f32)       # bias-corrected Adam step sizes of the D / G optimisers
    self.t_d = self.t_g = 0
    self.Wf, self.Wb = {}, {}
    self.refresh_weights()

  # -------------------------------------------------------------------------------------------
  def refresh_weights(self, which='GD'):
    """Derived (packed / TF32-rounded) filter copies for the tcgen05 path, after every Adam step of the
    named nets; refreshed in place so that captured graphs keep pointing at them."""
    P = self.P
    if 'G' in which:
      for i in range(1, 5):
        k = P['G/upconv_%d/W' % i]
        n = self.up[i].name
        self.Wf[n] = nets._pack_for_tc(self.up[i], k, self.gch[i - 1], self.Wf.get(n))
        self.Wb[n] = nets._pack_for_tc(self.up_b[i], k, self.gch[i], self.Wb.get(n))
    if 'D' in which:
      for i in range(4):
        k = P['D/conv_%d/W' % i]
        n = self.conv[i].name
        self.Wf[n] = nets._pack_for_tc(self.conv[i], k, self.dch[i], self.Wf.get(n))
        self.Wb[n] = nets._pack_for_tc(self.conv_t[i], k, self.dch[i + 1], self.Wb.get(n))

  def _replay(self, key, fn):
    """Run `fn` (device work on static buffers only): eagerly the first time, then captured once and
    replayed.  Host-side state (step counters, RNG draws, input staging) stays outside `fn`."""
    if not self.use_graphs:
      return fn()
    g = self._graphs.get(key)
    if g is not None:
      self.replayed_launches += self._graph_launches[key]
      return g.replay()
    fn()
    torch.cuda.synchronize()
    n0 = N.launch_count()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
      fn()
    self._graph_launches[key] = N.launch_count() - n0
    self.capture_launches += self._graph_launches[key]    # counted by the library, but only recorded
    self._graphs[key] = g

  def _w(self, table, L, key):
    w = table.get(L.name)
    return w if w is not None else self.P[key]

  # -------------------------------------------------------------------------------------------
  # primitives
  # -------------------------------------------------------------------------------------------
  def _gemm(self, a, lda, b, ldb, c, ldc, M, Nn, K, ta=0, tb=0, acc=0, bias=None):
    N.call('advoc_gemm_f32', _ptr(a), lda, _ptr(b), ldb, _ptr(c), ldc, M, Nn, K, ta, tb, acc, _ptr(bias),
           _stream())

  def _bn_fwd(self, x, stats, gamma, beta, act, y):
    Cc = x.shape[-1]
    pixels = x.numel() // Cc
    stats.zero_()
    N.call('advoc_bn_stats', _ptr(x), Cc, pixels, Cc, _ptr(stats), _stream())
    N.call('advoc_bn_apply', _ptr(x), Cc, pixels, Cc, _ptr(stats), _ptr(gamma), _ptr(beta), BN_EPS, act, 0.2,
           _ptr(y), Cc, self.rnd, _stream())

  def _bn_track(self, x, stats, moving):
    Cc = x.shape[-1]
    N.call('advoc_bn_moving_update', _ptr(stats), x.numel() // Cc, Cc, self.bn_momentum, _ptr(moving[0]),
           _ptr(moving[1]), _stream())

  def _bn_infer(self, x, moving, gamma, beta, act, y):
    Cc = x.shape[-1]
    N.call('advoc_bn_inference', _ptr(x), Cc, x.numel() // Cc, Cc, _ptr(moving[0]), _ptr(moving[1]), _ptr(gamma),
           _ptr(beta), BN_EPS, act, 0.2, _ptr(y), Cc, self.rnd, _stream())

  def moving_averages(self):
    """{TF variable name: tensor} of the generator's batch-norm moving statistics."""
    out = {}
    for i in range(4):
      out[G_BN[i] + '/moving_mean'], out[G_BN[i] + '/moving_variance'] = self.g_moving[i]
    return out

  def load_moving_averages(self, moving):
    for i in range(4):
      self.g_moving[i][0].copy_(moving[G_BN[i] + '/moving_mean'])
      self.g_moving[i][1].copy_(moving[G_BN[i] + '/moving_variance'])

  def _bn_bwd(self, dy, y, x, stats, gamma, act, red, dx, bn_name, param_grads):
    Cc = x.shape[-1]
    pixels = x.numel() // Cc
    red.zero_()
    N.call('advoc_bn_backward', _ptr(dy), Cc, _ptr(y), Cc, _ptr(x), Cc, pixels, Cc, _ptr(stats), _ptr(gamma),
           BN_EPS, act, 0.2, _ptr(red), _ptr(dx), Cc, self.rnd, _stream())
    if param_grads:
      self.flat.G[bn_name + '/beta'].add_(red[:Cc])
      self.flat.G[bn_name + '/gamma'].add_(red[Cc:])

  def _conv(self, L, x, ldx, w, ep):
    fn = 'advoc_conv2d_fwd' if L.kind == 'conv' else 'advoc_conv2d_transpose_fwd'
    N.call(fn, C.byref(L.desc), _ptr(x), ldx, _ptr(w), C.byref(ep), _stream())

  def _wgrad(self, desc, big, ld_big, small, ld_small, name):
    N.call('advoc_conv2d_wgrad', C.byref(desc), _ptr(big), ld_big, _ptr(small), ld_small,
           _ptr(self.flat.G[name + '/W']), _stream())

  def _bgrad(self, dy, channels, name):
    N.call('advoc_bias_grad', _ptr(dy), channels, dy.numel() // channels, channels,
           _ptr(self.flat.G[name + '/b']), _stream())

  # -------------------------------------------------------------------------------------------
  # generator  (conv2d.py:95-150)
  # -------------------------------------------------------------------------------------------
  def generate(self, z, training=True):
    """z [B, 100] on the device -> G_z [B, 64, 80, 1] in (-1, 1).  training=True normalises with batch
    statistics and advances the moving averages (every evaluation of G_z in the reference's train graph
    runs the update ops: conv2d.py:143-148); training=False normalises with the moving averages."""
    P, B = self.P, self.B
    if z is not self.z:
      self.z.copy_(z)
    n0 = 4 * 5 * self.gch[0]
    self._gemm(self.z, Z_DIM, P['G/z_proj/W'], n0, self.gX[0], n0, B, n0, Z_DIM, bias=P['G/z_proj/b'])

    def bn(i):
      gamma, beta = P[G_BN[i] + '/gamma'], P[G_BN[i] + '/beta']
      if training:
        self._bn_fwd(self.gX[i], self.gstats[i], gamma, beta, N.ACT_RELU, self.gY[i])
        self._bn_track(self.gX[i], self.gstats[i], self.g_moving[i])
      else:
        self._bn_infer(self.gX[i], self.g_moving[i], gamma, beta, N.ACT_RELU, self.gY[i])

    bn(0)
    for i in range(1, 5):
      L = self.up[i]
      w = self._w(self.Wf, L, L.name + '/W')
      if i < 4:
        ep = _epilogue(P[L.name + '/b'], self.gX[i], self.gch[i], 0, N.ACT_NONE)
        self._conv(L, self.gY[i - 1], self.gch[i - 1], w, ep)
        bn(i)
      else:
        ep = _epilogue(P[L.name + '/b'], self.G_z, 1, 0, N.ACT_TANH)
        self._conv(L, self.gY[3], self.gch[3], w, ep)
    return self.G_z

  def _g_backward(self):
    """dG_z holds d loss / d G_z."""
    P, B = self.P, self.B
    N.call('advoc_tanh_backward', _ptr(self.dG_z), _ptr(self.G_z), _ptr(self.dX4), self.dX4.numel(), _stream())
    dX = self.dX4
    for i in range(4, 0, -1):
      L = self.up_b[i]
      c_big, c_small = self.gch[i], self.gch[i - 1]
      self._wgrad(L.desc, dX, c_big, self.gY[i - 1], c_small, L.name)
      self._bgrad(dX, c_big, L.name)
      ep = _epilogue(None, self.gdY[i - 1], c_small, 0, N.ACT_NONE)
      self._conv(L, dX, c_big, self._w(self.Wb, L, L.name + '/W'), ep)
      self._bn_bwd(self.gdY[i - 1], self.gY[i - 1], self.gX[i - 1], self.gstats[i - 1], P[G_BN[i - 1] + '/gamma'],
                   N.ACT_RELU, self.gred[i - 1], self.gdX[i - 1], G_BN[i - 1], True)
      dX = self.gdX[i - 1]
    n0 = 4 * 5 * self.gch[0]
    G = self.flat.G
    # dW = z^T dX0 ; db = 1^T dX0
    self._gemm(self.z, Z_DIM, dX, n0, G['G/z_proj/W'], n0, Z_DIM, n0, B, ta=1, acc=1)
    self._gemm(self.ones, B, dX, n0, G['G/z_proj/b'], n0, 1, n0, B, acc=1)

  # -------------------------------------------------------------------------------------------
  # discriminator  (conv2d.py:166-219)
  # -------------------------------------------------------------------------------------------
  def discriminate(self, x, bufs):
    """x [B, 64, 80, 1] -> logits [B] (kept in bufs.logits)."""
    P, B = self.P, self.B
    L = self.conv[0]
    ep = _epilogue(P[L.name + '/b'], bufs.Y[0], self.dch[1], 0, N.ACT_LRELU, round_tf32=self.rnd)
    self._conv(L, x, 1, self._w(self.Wf, L, L.name + '/W'), ep)
    for i in range(1, 4):
      L = self.conv[i]
      ep = _epilogue(P[L.name + '/b'], bufs.X[i], self.dch[i + 1], 0, N.ACT_NONE)
      self._conv(L, bufs.Y[i - 1], self.dch[i], self._w(self.Wf, L, L.name + '/W'), ep)
      self._bn_fwd(bufs.X[i], bufs.stats[i], P[D_BN[i - 1] + '/gamma'], P[D_BN[i - 1] + '/beta'], N.ACT_LRELU,
                   bufs.Y[i])
    k = 4 * 5 * self.dch[4]
    self._gemm(bufs.Y[3], k, P['D/out/W'], 1, bufs.logits, 1, B, 1, k, bias=P['D/out/b'])
    return bufs.logits

  def _d_backward(self, x, bufs, param_grads, input_grad, dst=None):
    """bufs.dlogits holds d loss / d logits.  The input gradient goes to `dst` (default dG_z)."""
    P, B, G = self.P, self.B, self.flat.G
    k = 4 * 5 * self.dch[4]
    if param_grads:
      self._gemm(bufs.Y[3], k, bufs.dlogits, 1, G['D/out/W'], 1, k, 1, B, ta=1, acc=1)
      self._gemm(self.ones, B, bufs.dlogits, 1, G['D/out/b'], 1, 1, 1, B, acc=1)
    # dY3 = dlogits (x) W^T
    self._gemm(bufs.dlogits, 1, P['D/out/W'], 1, bufs.dY[3], k, B, k, 1, tb=1)
    for i in range(3, 0, -1):
      L = self.conv[i]
      self._bn_bwd(bufs.dY[i], bufs.Y[i], bufs.X[i], bufs.stats[i], P[D_BN[i - 1] + '/gamma'], N.ACT_LRELU,
                   bufs.red[i], bufs.dX[i], D_BN[i - 1], param_grads)
      if param_grads:
        self._wgrad(L.desc, bufs.Y[i - 1], self.dch[i], bufs.dX[i], self.dch[i + 1], L.name)
        self._bgrad(bufs.dX[i], self.dch[i + 1], L.name)
      Lt = self.conv_t[i]
      if i > 1:
        ep = _epilogue(None, bufs.dY[i - 1], self.dch[i], 0, N.ACT_NONE)
      else:   # conv_0 has no batch norm: fold its leaky-ReLU derivative into this epilogue
        ep = _epilogue(None, bufs.dX[0], self.dch[1], 0, N.ACT_NONE, gate=bufs.Y[0], ld_gate=self.dch[1],
                       gate_act=N.ACT_LRELU, round_tf32=self.rnd)
      self._conv(Lt, bufs.dX[i], self.dch[i + 1], self._w(self.Wb, Lt, Lt.name + '/W'), ep)
    L = self.conv[0]
    if param_grads:
      self._wgrad(L.desc, x, 1, bufs.dX[0], self.dch[1], L.name)
      self._bgrad(bufs.dX[0], self.dch[1], L.name)
    if input_grad:
      Lt = self.conv_t[0]
      ep = _epilogue(None, self.dG_z if dst is None else dst, 1, 0, N.ACT_NONE)
      self._conv(Lt, bufs.dX[0], self.dch[1], self._w(self.Wb, Lt, Lt.name + '/W'), ep)

  # -------------------------------------------------------------------------------------------
  # gradient penalty (train.py:99-109): loss term and its gradient w.r.t. the critic's parameters
  # -------------------------------------------------------------------------------------------
  def _gp_backward(self, alpha=None):
    """x_real and G_z are in place.  Adds LAMBDA * mean((||grad_xhat D(xhat)||_2 - 1)^2) to
    losses[0] and its parameter gradient to the critic's slice of the flat gradient buffer."""
    P, B, G, I = self.P, self.B, self.flat.G, self.interp
    ch = self.dch
    if alpha is None:
      alpha = self.alpha_buf.uniform_(0., 1., generator=self._alpha_gen)
    torch.lerp(self.x_real, self.G_z, alpha, out=self.x_hat)   # x + alpha (G_z - x)
    # first order: D(xhat) and g = d sum_b D(xhat)_b / d xhat
    self.discriminate(self.x_hat, I)
    I.dlogits.fill_(1.0)
    self._d_backward(self.x_hat, I, False, True, dst=self.g_hat)
    n = self.g_hat.numel() // B
    N.call('advoc_gp_seed', _ptr(self.g_hat), B, n, self.gp_lambda, _ptr(self.losses), _ptr(self.u_hat), self.rnd,
           _stream())
    # ---- up sweep: adjoint of the first backward pass (u = d penalty / d g)
    L = self.conv[0]
    ep = _epilogue(None, I.vy[0], ch[1], 0, N.ACT_NONE, gate=I.Y[0], ld_gate=ch[1], gate_act=N.ACT_LRELU,
                   round_tf32=self.rnd)
    self._conv(L, self.u_hat, 1, self._w(self.Wf, L, L.name + '/W'), ep)        # adjoint of dY_0 (masked)
    self._wgrad(L.desc, self.u_hat, 1, I.dX[0], ch[1], L.name)
    for i in range(1, 4):
      L = self.conv[i]
      Cc = ch[i + 1]
      pixels = I.X[i].numel() // Cc
      ep = _epilogue(None, I.v[i], Cc, 0, N.ACT_NONE)
      self._conv(L, I.vy[i - 1], ch[i], self._w(self.Wf, L, L.name + '/W'), ep)   # adjoint of dX_i
      self._wgrad(L.desc, I.vy[i - 1], ch[i], I.dX[i], Cc, L.name)
      I.sums[i].zero_()
      gamma = P[D_BN[i - 1] + '/gamma']
      N.call('advoc_bn_gp', _ptr(I.v[i]), _ptr(I.dY[i]), _ptr(I.Y[i]), _ptr(I.X[i]), pixels, Cc, _ptr(I.stats[i]),
             _ptr(gamma), BN_EPS, 0.2, _ptr(I.sums[i]), _ptr(I.vy[i]), _ptr(I.xb[i]), self.rnd, _stream())
      # d penalty / d gamma = S / gamma = 1/sigma (sum v a - sum v sum a / M - sum v xhat sum a xhat / M)
      st = I.stats[i]
      mean = st[:Cc] / pixels
      invstd = torch.rsqrt((st[Cc:] / pixels - mean * mean).clamp_min(0) + BN_EPS)
      sv, sa, sva, svx, sax = I.sums[i].view(5, Cc)
      G[D_BN[i - 1] + '/gamma'].add_(invstd * (sva - sv * sa / pixels - svx * sax / pixels))
    k = 4 * 5 * ch[4]
    self._gemm(self.ones, B, I.vy[3], k, G['D/out/W'], k, 1, k, B, acc=1)       # dY_3 = 1 (x) w_out
    # ---- down sweep: ordinary backward pass seeded with the adjoints of the forward X_i
    for i in range(3, 0, -1):
      L = self.conv[i]
      Cc = ch[i + 1]
      self._wgrad(L.desc, I.Y[i - 1], ch[i], I.xb[i], Cc, L.name)
      self._bgrad(I.xb[i], Cc, L.name)
      Lt = self.conv_t[i]
      if i > 1:
        ep = _epilogue(None, I.tmp[i - 1], ch[i], 0, N.ACT_NONE)
        self._conv(Lt, I.xb[i], Cc, self._w(self.Wb, Lt, Lt.name + '/W'), ep)
        # through lrelu + batch norm of layer i-1; its input adjoint joins the second-order term
        self._bn_bwd(I.tmp[i - 1], I.Y[i - 1], I.X[i - 1], I.stats[i - 1], P[D_BN[i - 2] + '/gamma'], N.ACT_LRELU,
                     I.red[i - 1], I.v[i - 1], D_BN[i - 2], True)
        I.xb[i - 1].add_(I.v[i - 1])
      else:
        ep = _epilogue(None, I.xb[0], ch[1], 0, N.ACT_NONE, gate=I.Y[0], ld_gate=ch[1], gate_act=N.ACT_LRELU)
        self._conv(Lt, I.xb[1], Cc, self._w(self.Wb, Lt, Lt.name + '/W'), ep)
    L = self.conv[0]
    self._wgrad(L.desc, self.x_hat, 1, I.xb[0], ch[1], L.name)
    self._bgrad(I.xb[0], ch[1], L.name)

  # -------------------------------------------------------------------------------------------
  # optimiser + collective
  # -------------------------------------------------------------------------------------------
  def _finish(self, lo, hi, t, apply, which):
    from advoc_b200 import dist as D
    D.allreduce_sum_(self.flat.g, lo, hi, self.pg, self.world)
    if not apply:
      return
    # bias-corrected step size of tf.train.AdamOptimizer, kept on the device so that the captured update
    # does not freeze t
    slot = 0 if which == 'D' else 1
    self.lr_t[slot:slot + 1].fill_(self.lr * (1.0 - self.b2 ** t) ** 0.5 / (1.0 - self.b1 ** t))

    def update():
      f = self.flat
      o = lambda t_: C.c_void_p(t_.data_ptr() + 4 * lo)
      N.call('advoc_adam_tf_step_dev', o(f.p), o(f.g), o(f.m), o(f.v), hi - lo,
             C.c_void_p(self.lr_t.data_ptr() + 4 * slot), self.b1, self.b2, self.eps, 1.0 / self.world, _stream())
      self.refresh_weights(which)
    self._replay('adam_' + which, update)

  def _d_compute(self):
    lo, hi = self.flat.dis_range()
    self.flat.g[lo:hi].zero_()
    G_z = self.generate(self.z)
    self.discriminate(self.x_real, self.real)
    self.discriminate(G_z, self.fake)
    N.call('advoc_gan_logit_loss', _ptr(self.real.logits), _ptr(self.fake.logits), self.B,
           0 if self.loss_kind == 'dcgan' else 2, _ptr(self.losses), _ptr(self.real.dlogits),
           _ptr(self.fake.dlogits), _stream())
    self._d_backward(self.x_real, self.real, True, False)
    self._d_backward(G_z, self.fake, True, False)
    if self.loss_kind == 'wgangp':
      self._gp_backward(self.alpha_buf)

  def d_step(self, x, z, apply=True, alpha=None):
    """`D_train_op` on one minibatch (train.py:139,151): x [B,64,80,1] in [-1,1], z [B,100];
    `alpha` [B,1,1,1] fixes the interpolation draw of train.py:100 (tests)."""
    lo, hi = self.flat.dis_range()
    self.x_real.copy_(x)
    self.z.copy_(z)
    if self.loss_kind == 'wgangp':
      if alpha is None:
        self.alpha_buf.uniform_(0., 1., generator=self._alpha_gen)
      else:
        self.alpha_buf.copy_(alpha.reshape(self.alpha_buf.shape))
    self._replay('d', self._d_compute)
    if apply:
      self.t_d += 1
    self._finish(lo, hi, self.t_d, apply, 'D')

  def _g_compute(self):
    lo, hi = self.flat.gen_range()
    self.flat.g[lo:hi].zero_()
    G_z = self.generate(self.z)
    self.discriminate(G_z, self.fake)
    N.call('advoc_gan_logit_loss', None, _ptr(self.fake.logits), self.B, 1 if self.loss_kind == 'dcgan' else 3,
           C.c_void_p(self.losses.data_ptr() + 4), None, _ptr(self.fake.dlogits), _stream())
    self._d_backward(G_z, self.fake, False, True)
    self._g_backward()

  def g_step(self, z, apply=True):
    """`G_train_op` on one draw of z (train.py:137-138,153)."""
    lo, hi = self.flat.gen_range()
    self.z.copy_(z)
    self._replay('g', self._g_compute)
    if apply:
      self.t_g += 1
    self._finish(lo, hi, self.t_g, apply, 'G')
    return self.t_g

  def train_loop(self, batches, zs_d, z_g):
    """One outer iteration of train.py:149-153: `num_disc_updates_per_genr` D steps (1 for dcgan,
    5 for wgangp), each on a fresh minibatch and a fresh z, then one G step."""
    for x, z in zip(batches, zs_d):
      self.d_step(x, z)
    return self.g_step(z_g)

  def loss_values(self):
    v = self.losses.tolist()
    N.raise_if_aborted('MelspecGAN')
    return v[0], v[1]


# ---------------------------------------------------------------------------------------------
# Callable mirrors of the reference classes (models/melspecgan/conv2d.py:82-95,153-166): same
# names, constructor arguments and call signature; tensors are CUDA float32 in the reference's
# shapes.  The generator's `training=False` graph (infer / incept, train.py:163,192) normalises with the
# moving averages the engine tracks; the discriminator is never built with training=False by the reference.
# ---------------------------------------------------------------------------------------------
class _Net(object):
  def __init__(self, dim=64, kernel_len=5, batchnorm=True, params=None):
    if kernel_len != 5 or not batchnorm:
      raise NotImplementedError()
    self.dim, self.kernel_len, self.stride, self.batchnorm = dim, kernel_len, 2, batchnorm
    self.params = params
    self._eng = None

  def _engine(self, batch):
    if self._eng is None or self._eng.B != batch:
      if self.params is None:
        self.params = init_params(dim=self.dim)
      self._eng = MelspecGAN(self.params, batch, dim=self.dim)
    return self._eng


class MelspecGANGenerator(_Net):
  def __init__(self, dim=64, kernel_len=5, batchnorm=True, params=None, moving=None):
    super(MelspecGANGenerator, self).__init__(dim, kernel_len, batchnorm, params)
    self.moving = moving      # {'G/batch_normalization*/moving_{mean,variance}': tensor}, e.g. from a checkpoint

  def __call__(self, z, training=False):
    fresh = self._eng is None or self._eng.B != z.shape[0]
    eng = self._engine(z.shape[0])
    if fresh and self.moving is not None:
      eng.load_moving_averages(self.moving)
    out = eng.generate(z, training=training)
    self.moving = eng.moving_averages()
    return out


class MelspecGANDiscriminator(_Net):
  def __call__(self, x, training=False):
    if not training:
      raise NotImplementedError('the discriminator\'s moving-average batch norm is not tracked')
    eng = self._engine(x.shape[0])
    return eng.discriminate(x.contiguous(), eng.real)
