"""Execution engine for the AdVoc conv stacks on B200 (generator U-Net + PatchGAN discriminator).

Host side only: shape arithmetic (TensorFlow SAME/VALID rules), persistent NHWC activation
buffers in HBM, and the per-layer launches into libadvoc_b200.so.  Everything the reference
does between two convolutions -- bias, lrelu/relu, the skip concat, the `[:, :, :-1, :]` crop and
dropout -- is folded into the producing kernel's epilogue (dual write into the next encoder's
input and into the channel slice of the decoder concat buffer), so no element-wise or copy
kernel runs between layers.

reference: models/advoc/advoc_model.py:25-69 (layer primitives), :75-166 (generator),
:168-204 (discriminator); models/advoc/advoc_model_small.py (truncated variant).

Parameters are a dict keyed by the TF variable names
(`generator/encoder_N/conv2d/{kernel,bias}`, `generator/decoder_N/conv2d_transpose/{kernel,bias}`,
`discriminator/layer_N/conv2d/{kernel,bias}`) holding CUDA float32 tensors in the TF layouts
(conv HWIO, conv_transpose HWOI), so reference checkpoints map 1:1.
"""
import ctypes as C

import torch

from advoc_b200 import _native as N


def _ptr(t):
  return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def same_pads(n, k, s):
  """TF 'SAME': out = ceil(n/s), pad_total = max((out-1)s + k - n, 0), before = total // 2."""
  out = -(-n // s)
  total = max((out - 1) * s + k - n, 0)
  return out, total // 2, total - total // 2


class GenSpec(object):
  """Static geometry of a generator (advoc_model.py:91-158)."""

  def __init__(self, ngf, n_enc, dropout_decoders, subseq_len=256, nbins=513):
    self.ngf, self.n_enc = ngf, n_enc
    self.dropout_decoders = tuple(dropout_decoders)
    mult = [1, 2, 4, 8, 8, 8, 8, 8]
    self.enc_ch = [ngf * m for m in mult[:n_enc]]
    dmult = {8: 8, 7: 8, 6: 8, 5: 8, 4: 4, 3: 2, 2: 1}
    # the reference keeps the LAST (n_enc - 1) decoder specs (advoc_model_small.py:134)
    self.dec_ch = {k: ngf * dmult[k] for k in range(n_enc, 1, -1)}
    self.dec_ch[1] = 1
    # spatial sizes: index 0 = network input, i = output of encoder_i
    self.H, self.W, self.sh = [subseq_len], [nbins], [None]
    n_time = subseq_len
    self.n_stride1 = 0
    for i in range(1, n_enc + 1):
      if i == 1 or n_time > 1:
        sh = 2
        n_time //= 2
      else:
        sh = 1
        self.n_stride1 += 1
      self.sh.append(sh)
      self.H.append(-(-self.H[-1] // sh))
      self.W.append(-(-self.W[-1] // 2))


class _Conv(object):
  """One convolution launch: geometry + operands; kind 'conv' or 'deconv'."""

  def __init__(self, name, kind, desc):
    self.name, self.kind, self.desc = name, kind, desc
    self.n_out = 1

  def run(self, x, ldx, w, ep):
    fn = 'advoc_conv2d_fwd' if self.kind == 'conv' else 'advoc_conv2d_transpose_fwd'
    self.n_out = 2 if ep.d_out1 else 1
    self.store_w = ep.store_w
    self.ldx = ldx
    N.call(fn, C.byref(self.desc), _ptr(x), ldx, _ptr(w), C.byref(ep), _stream())

  # -- accounting used by bench.py (dense MACs x 2) --------------
  def path(self, ldx=None):
    """advoc_conv2d_path: MATH_FP32 (CUDA cores), MATH_TF32 or MATH_F16 (tcgen05)."""
    ldx = self.ldx if ldx is None else ldx
    return N.lib().advoc_conv2d_path(C.byref(self.desc), ldx, 1 if self.kind == 'deconv' else 0)

  def uses_tensor_cores(self, ldx=None):
    return self.path(ldx) != N.MATH_FP32

  def half_operands(self):
    return self.desc.math == N.MATH_F16

  def tile_n(self):
    """BN of the tcgen05 instantiation this layer launches (0 on the CUDA-core kernels)."""
    return N.lib().advoc_conv2d_tile_n(C.byref(self.desc), self.ldx, 1 if self.kind == 'deconv' else 0,
                                       getattr(self, 'store_w', 0) or 0)

  def kernel_family(self):
    """Kernel that runs this layer (bench.py attributes time by it)."""
    k = N.lib().advoc_conv2d_kernel(C.byref(self.desc), self.ldx, 1 if self.kind == 'deconv' else 0,
                                    getattr(self, 'store_w', 0) or 0)
    if k == 4:
      return 'conv_one_in_tc_kernel/' + self.kind
    if k == 3:
      return 'deconv_one_tc_kernel/' + self.kind
    if k == 2:
      return 'conv_p2d_kernel/' + self.kind
    if k == 1:
      return 'conv_tc_kernel/' + self.kind
    d = self.desc
    if self.kind == 'conv' and d.Cin <= 2:
      return 'thin_conv'
    if self.kind == 'deconv' and d.Cin == 1:
      return 'to_one_deconv'
    return 'simt_' + self.kind

  def flops(self):
    """Dense MACs x 2 of the layer; a layer run on a padded geometry (the pixel-pair
    encoder_2 of the fp16 generator) reports the flops of the layer it implements."""
    if getattr(self, 'dense_flops', None) is not None:
      return self.dense_flops
    d = self.desc
    return 2.0 * d.N * d.Ho * d.Wo * d.kh * d.kw * d.Cin * d.Cout

  def algorithmic_bytes(self):
    """Bytes the layer must move once: input + filter + stored output(s), at the element sizes of
    the buffers it actually ran on (`in_bytes` / `out_bytes`, recorded by the engine)."""
    d = self.desc
    big, small = d.N * d.H * d.W * d.Cin, d.N * d.Ho * d.Wo * d.Cout
    w = d.kh * d.kw * d.Cin * d.Cout
    ib, ob = getattr(self, 'in_bytes', 4), getattr(self, 'out_bytes', 4)
    if self.kind == 'conv':
      return ib * (big + w) + ob * small * self.n_out
    ws = self.store_w or d.W
    return ib * (small + w) + ob * d.N * d.H * ws * d.Cin


def _desc(n, h, w, cin, cout, sh, sw, pt, pl, ho, wo, math):
  return N.ConvDesc(n, h, w, cin, cout, 4, 4, sh, sw, pt, pl, ho, wo, math)


def _epilogue(bias, out0, ld0, coff0, act0, out1=None, ld1=0, coff1=0, act1=N.ACT_NONE,
              store_w=0, mask=None, keep_prob=1.0, seed=0, round_tf32=0, alpha=0.2,
              accumulate=0, gate=None, ld_gate=0, coff_gate=0, gate_act=N.ACT_NONE, gate_split=0,
              gate_scale0=1.0, gate_scale1=1.0, seed_dev=None, row_pad0=0):
  """advoc_epilogue; the output dtypes follow the destination tensors (float32 / float16)."""
  h0 = N.DT_F16 if (out0 is not None and out0.dtype == torch.float16) else N.DT_F32
  h1 = N.DT_F16 if (out1 is not None and out1.dtype == torch.float16) else N.DT_F32
  return N.Epilogue(_ptr(bias), act0, act1, alpha, _ptr(out0), ld0, coff0, _ptr(out1), ld1, coff1,
                    store_w, _ptr(mask), keep_prob, seed, round_tf32, accumulate, _ptr(gate),
                    ld_gate, coff_gate, gate_act, gate_split, gate_scale0, gate_scale1, _ptr(seed_dev),
                    h0, h1, row_pad0, 0)


def _pack_for_tc(L, kernel, ldx, out=None):
  """Derived filter copy for the tcgen05 path (None if the layer runs on CUDA cores):
  conv HWIO -> K-major [tap][Cout][Cin]; conv_transpose HWOI is already K-major.  Values are
  rounded to TF32 (round-to-nearest) because the tensor core would otherwise truncate.
  `out`: an earlier result to refresh in place (fixed addresses for captured graphs)."""
  if not L.uses_tensor_cores(ldx):
    return None
  kh, kw, a, b = kernel.shape
  half = L.half_operands()   # ADVOC_MATH_F16 layers read an fp16 copy (same 10-bit mantissa as TF32)
  if out is None:
    out = torch.empty((kh * kw, b, a) if L.kind == 'conv' else (kh * kw, a, b),
                      dtype=torch.float16 if half else torch.float32, device=kernel.device)
  N.call('advoc_pack_filter', _ptr(kernel), _ptr(out), kh * kw, a, b, 1 if L.kind == 'conv' else 0,
         2 if half else 1, _stream())
  return out


def _pair_filter(kernel):
  """HWIO filter [4, 4, 32, Cout] of a k4 s2 convolution along W -> [4, 3, 64, Cout] of the equivalent k3 s1
  convolution over PIXEL PAIRS (channel index = parity * 32 + c, pair j = pixels 2j, 2j+1; SAME padding (1, 2)
  of the odd-width input becomes one zero pair on each side).  Output column ow reads input columns
  2ow-1 .. 2ow+2 = pair ow-1 (odd pixel: tap 0), pair ow (taps 1, 2), pair ow+1 (even pixel: tap 3)."""
  kh, kw, cin, cout = kernel.shape
  assert kw == 4 and cin == 32
  out = torch.zeros((kh, 3, 2 * cin, cout), dtype=kernel.dtype, device=kernel.device)
  out[:, 0, cin:] = kernel[:, 0]
  out[:, 1, :cin] = kernel[:, 1]
  out[:, 1, cin:] = kernel[:, 2]
  out[:, 2, :cin] = kernel[:, 3]
  return out


class Generator(object):
  """U-Net generator forward on persistent buffers.  x [B,T,513,1] -> [B,T,513,1].

  math = MATH_F16 (inference): every layer whose contraction side has a multiple of 64 channels reads
  fp16 activations and filters (tcgen05 kind::f16; half the HBM / L2 / shared-memory operand bytes of
  the TF32 path at the same 10-bit mantissa) and the buffers between such layers are stored as fp16 by
  the producing epilogue; the other layers (encoder_1 on CUDA cores; a 32-channel encoder_2 input) stay
  fp32 / TF32.  The network input and output are fp32 either way."""

  def __init__(self, spec, params, batch, math=N.MATH_AUTO, device=None):
    self.spec, self.P, self.B, self.math = spec, params, batch, math
    dev = device or torch.device('cuda', torch.cuda.current_device())
    self.dev = dev
    s = spec
    n = s.n_enc
    self.Dk = {k: (s.dec_ch[k + 1] if k < n else 0) for k in range(1, n + 1)}
    self.enc, self.dec = {}, {}
    base_math = N.MATH_AUTO if math == N.MATH_F16 else math

    def pick(kind, name, ldx, mk):
      """The layer with fp16 operands when asked for and eligible, else in the base math mode."""
      if math == N.MATH_F16:
        L = _Conv(name, kind, mk(N.MATH_F16))
        if L.path(ldx) == N.MATH_F16:
          return L
      return _Conv(name, kind, mk(base_math))

    for i in range(1, n + 1):
      cin = 1 if i == 1 else s.enc_ch[i - 2]
      ho, pt, _ = same_pads(s.H[i - 1], 4, s.sh[i])
      wo, pl, _ = same_pads(s.W[i - 1], 4, 2)
      assert ho == s.H[i] and wo == s.W[i]
      self.enc[i] = pick('conv', 'generator/encoder_%d/conv2d' % i, max(cin, 1),
                         lambda m, i=i, cin=cin, ho=ho, wo=wo, pt=pt, pl=pl: _desc(
                             batch, s.H[i - 1], s.W[i - 1], cin, s.enc_ch[i - 1], s.sh[i], 2, pt, pl, ho, wo, m))
    for j, k in enumerate(range(n, 0, -1)):
      sh = 1 if j < s.n_stride1 else 2
      cin = self.Dk[k] + s.enc_ch[k - 1]
      # geometry of the forward conv this deconv is the input-gradient of: big side = output
      self.dec[k] = pick('deconv', 'generator/decoder_%d/conv2d_transpose' % k, cin,
                         lambda m, k=k, sh=sh, cin=cin: _desc(batch, s.H[k] * sh, s.W[k] * 2, s.dec_ch[k], cin, sh, 2,
                                                              1, 1, s.H[k], s.W[k], m))
      assert s.H[k] * sh == s.H[k - 1] and s.W[k] * 2 - 1 == s.W[k - 1]
    # AdVoc-small: encoder_2 reads 32 channels, half a 128-byte fp16 k-block.  Store encoder_1's output
    # as fp16 with every row padded by one zero pixel (257 -> 258) and let encoder_2 read PAIRS of pixels
    # as 64-channel vectors: the k4 s2 convolution along W becomes a k3 s1 convolution over 129 pixel
    # pairs whose filter holds the four real taps and zeros (`_pair_filter`): 1.5x the MACs, but fp16
    # MACs on fp16 bytes instead of TF32 ones on fp32 bytes.
    self.pair2 = None
    if (math == N.MATH_F16 and n >= 2 and s.enc_ch[0] == 32 and s.W[1] % 2 == 1 and
        not self.enc[2].half_operands()):
      wp = (s.W[1] + 1) // 2
      L = _Conv('generator/encoder_2/conv2d', 'conv',
                _desc(batch, s.H[1], wp, 64, s.enc_ch[1], s.sh[2], 1, self.enc[2].desc.pad_t, 1, s.H[2], wp,
                      N.MATH_F16))
      L.desc.kw = 3
      if wp == s.W[2] and L.path(64) == N.MATH_F16:
        self.pair2 = L
        L.dense_flops = self.enc[2].flops()     # accounted on the dense geometry (the zero taps are not work)
        self.enc[2] = L
    # buffer element types follow their (single) consumer
    dt = lambda L: torch.float16 if L.half_operands() else torch.float32
    # lrelu(encoder_i) for i < n: the next encoder's input
    self.E = {i: torch.empty((batch, s.H[i], s.W[i], s.enc_ch[i - 1]), dtype=dt(self.enc[i + 1]), device=dev)
              for i in range(1, n)}
    if self.pair2 is not None:
      self.E[1] = torch.zeros((batch, s.H[1], s.W[1] + 1, s.enc_ch[0]), dtype=torch.float16, device=dev)
    # decoder_k input = relu(concat(decoder_{k+1}[:, :, :-1], encoder_k))
    self.Cat = {k: torch.empty((batch, s.H[k], s.W[k], self.Dk[k] + s.enc_ch[k - 1]), dtype=dt(self.dec[k]),
                               device=dev)
                for k in range(1, n + 1)}
    self.out = torch.empty((batch, s.H[0], s.W[0], 1), dtype=torch.float32, device=dev)
    es = lambda t: t.element_size()
    for i in range(1, n + 1):
      self.enc[i].in_bytes = 4 if i == 1 else es(self.E[i - 1])
      self.enc[i].out_bytes = (es(self.E[i]) + es(self.Cat[i])) / 2.0 if i < n else es(self.Cat[i])
    for k in range(1, n + 1):
      self.dec[k].in_bytes = es(self.Cat[k])
      self.dec[k].out_bytes = es(self.Cat[k - 1]) if k > 1 else 4

  def _ldx(self, L):
    if L.kind == 'conv':
      return max(L.desc.Cin, 1)
    return L.desc.Cout

  def prepare(self):
    """Refresh the derived (packed / TF32-rounded) filter copies after the parameters changed
    and decide which stored activations must be TF32-rounded for their consumer."""
    old = getattr(self, 'Wp', {})   # refreshed in place: captured CUDA graphs keep pointing at these copies
    self.Wp = {}
    for L in list(self.enc.values()) + list(self.dec.values()):
      k = self.P[L.name + '/kernel']
      if L is self.pair2:
        k = _pair_filter(k)
      self.Wp[L.name] = _pack_for_tc(L, k, self._ldx(L), old.get(L.name))
    n = self.spec.n_enc
    tc = lambda L: self.Wp[L.name] is not None
    # encoder_i writes E[i] (read by encoder_{i+1}) and Cat[i] (read by decoder_i)
    self.round_enc = {i: int((i < n and tc(self.enc[i + 1])) or tc(self.dec[i])) for i in range(1, n + 1)}
    # decoder_k writes Cat[k-1] (read by decoder_{k-1})
    self.round_dec = {k: int(k > 1 and tc(self.dec[k - 1])) for k in range(1, n + 1)}
    return self

  def _w(self, L):
    if not hasattr(self, 'Wp'):
      self.prepare()
    w = self.Wp[L.name]
    return w if w is not None else self.P[L.name + '/kernel']

  def _run_layer(self, L, x, ldx, w, ep):
    L.run(x, ldx, w, ep)

  def dropout_shape(self, k):
    """Shape of the dropout mask of decoder_k: its stored (cropped) output."""
    s = self.spec
    return (self.B, s.H[k - 1], s.W[k - 1], s.dec_ch[k])

  def forward(self, x, out=None, out_ld=1, out_coff=0, dropout=None, seed=0, x_ld=1, seed_dev=None):
    """x f32 [B,T,513,1] contiguous.  dropout: None (off, parity mode) | 'rng' (counter-based
    generator keyed by `seed`, the reference's behaviour in every mode, advoc_model.py:144-149)
    | {decoder_index: uint8 mask tensor} (injected masks).  `seed_dev`: a one-element int64 CUDA
    tensor holding the step counter; the kernels then read the seed from it (a captured CUDA graph
    draws fresh masks on every replay; same masks as `seed` = the counter's value on the eager path).
    Returns the output buffer (`out` if given: written with pixel stride out_ld at channel out_coff)."""
    s, P, n = self.spec, self.P, self.spec.n_enc
    assert x.is_contiguous() and tuple(x.shape) == (self.B, s.H[0], s.W[0], x_ld), x.shape
    if not hasattr(self, 'Wp'):
      self.prepare()
    inp, ld = x, x_ld   # x_ld > 1: the input is channel 0 of a wider buffer
    for i in range(1, n + 1):
      L = self.enc[i]
      cat = self.Cat[i]
      if i < n:
        ep = _epilogue(P[L.name + '/bias'], self.E[i], s.enc_ch[i - 1], 0, N.ACT_LRELU,
                       cat, cat.shape[3], self.Dk[i], N.ACT_RELU, round_tf32=self.round_enc[i],
                       row_pad0=1 if (i == 1 and self.pair2 is not None) else 0)
      else:
        ep = _epilogue(P[L.name + '/bias'], cat, cat.shape[3], 0, N.ACT_RELU,
                       round_tf32=self.round_enc[i])
      self._run_layer(L, inp, ld, self._w(L), ep)
      if i < n:
        inp, ld = self.E[i], s.enc_ch[i - 1]
        if i == 1 and self.pair2 is not None:
          ld = 64       # pixel pairs
    dst = self.out if out is None else out
    for k in range(n, 0, -1):
      L = self.dec[k]
      cat = self.Cat[k]
      kw = {}
      if k in s.dropout_decoders and dropout is not None:
        kw['keep_prob'] = 0.5
        if dropout == 'rng' and seed_dev is not None:
          kw['seed'], kw['seed_dev'] = k, seed_dev
        elif dropout == 'rng':
          kw['seed'] = (seed * 0x9E3779B1 + k) & 0xFFFFFFFFFFFFFFFF
        else:
          kw['mask'] = dropout[k]
      if k > 1:
        nxt = self.Cat[k - 1]
        ep = _epilogue(P[L.name + '/bias'], nxt, nxt.shape[3], 0, N.ACT_RELU,
                       store_w=s.W[k - 1], round_tf32=self.round_dec[k], **kw)
      else:
        ld_o, co = (1, 0) if out is None else (out_ld, out_coff)
        ep = _epilogue(P[L.name + '/bias'], dst, ld_o, co, N.ACT_NONE, store_w=s.W[0], **kw)
      self._run_layer(L, cat, cat.shape[3], self._w(L), ep)
    return dst


class Discriminator(object):
  """PatchGAN forward.  in_cat f32 [B,T,513,2] (cond, target-or-generated) -> sigmoid map."""
  STRIDES = (2, 2, 2, 1, 1)

  def __init__(self, ndf, params, batch, math=N.MATH_AUTO, subseq_len=256, nbins=513, device=None):
    self.P, self.B = params, batch
    dev = device or torch.device('cuda', torch.cuda.current_device())
    chans = [2, ndf, ndf * 2, ndf * 4, ndf * 8, 1]
    self.chans = chans
    h, w = subseq_len, nbins
    self.layers, self.act = [], []
    for i, st in enumerate(self.STRIDES):
      ho, wo = (h + 2 - 4) // st + 1, (w + 2 - 4) // st + 1
      self.layers.append(_Conv('discriminator/layer_%d/conv2d' % (i + 1), 'conv',
                               _desc(batch, h, w, chans[i], chans[i + 1], st, st, 1, 1, ho, wo,
                                     math)))
      self.act.append(torch.empty((batch, ho, wo, chans[i + 1]), dtype=torch.float32, device=dev))
      h, w = ho, wo

  def prepare(self):
    old = getattr(self, 'Wp', {})   # refreshed in place (fixed addresses for captured graphs)
    self.Wp = {L.name: _pack_for_tc(L, self.P[L.name + '/kernel'], self.chans[i], old.get(L.name))
               for i, L in enumerate(self.layers)}
    self.round = [int(i + 1 < 5 and self.Wp[self.layers[i + 1].name] is not None) for i in range(5)]
    # layer_4's activation is the TMA-fed tensor-core operand of the head's filter gradient
    # (csrc/wgrad_thin_tc.cu), which would otherwise TRUNCATE it to tf32
    if self.layers[4].desc.math != N.MATH_FP32 and self.round[2]:
      self.round[3] = 1
    return self

  def _w(self, L):
    w = self.Wp[L.name]
    return w if w is not None else self.P[L.name + '/kernel']

  def forward(self, in_cat):
    assert in_cat.is_contiguous() and in_cat.shape[3] == 2 and in_cat.shape[0] == self.B
    if not hasattr(self, 'Wp'):
      self.prepare()
    x, ld = in_cat, 2
    for i, L in enumerate(self.layers):
      act = N.ACT_LRELU if i < 4 else N.ACT_SIGMOID
      ep = _epilogue(self.P[L.name + '/bias'], self.act[i], self.chans[i + 1], 0, act,
                     round_tf32=self.round[i])
      L.run(x, ld, self._w(L), ep)
      x, ld = self.act[i], self.chans[i + 1]
    return x


def init_params(ngf, ndf, n_enc, seed=0, device='cuda'):
  """N(0, 0.02) kernels and zero biases under the TF variable names
  (advoc_model.py:32,36,55; tf.layers default bias initialiser)."""
  g = torch.Generator(device='cpu').manual_seed(seed)
  spec = GenSpec(ngf, n_enc, ())
  P = {}

  def add(name, shape, nbias):
    P[name + '/kernel'] = (torch.randn(shape, generator=g) * 0.02).to(device)
    P[name + '/bias'] = torch.zeros(nbias, device=device)

  cin = 1
  for i, c in enumerate(spec.enc_ch):
    add('generator/encoder_%d/conv2d' % (i + 1), (4, 4, cin, c), c)
    cin = c
  for k in range(n_enc, 0, -1):
    cin = spec.enc_ch[k - 1] + (spec.dec_ch[k + 1] if k < n_enc else 0)
    add('generator/decoder_%d/conv2d_transpose' % k, (4, 4, spec.dec_ch[k], cin), spec.dec_ch[k])
  cin = 2
  for i, c in enumerate([ndf, ndf * 2, ndf * 4, ndf * 8, 1]):
    add('discriminator/layer_%d/conv2d' % (i + 1), (4, 4, cin, c), c)
    cin = c
  return P
