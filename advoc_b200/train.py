"""G + D train step of the adversarial vocoder on B200 (forward, backward, TF1-Adam), with
data-parallel gradient all-reduce over NCCL.

reference: models/advoc/advoc_model.py:206-257 (`__call__`: losses, variable partition, two Adam
optimisers) and :285-289 (`train_loop`: D update on one minibatch, then G update on the next,
through the already-updated D).  The reference gets its backward pass from TF autodiff; here it
is laid out explicitly on the same persistent NHWC buffers as the forward pass:

  * input gradients reuse the forward tcgen05 kernels (conv dgrad == transposed conv with the
    conv's HWIO filter; deconv dgrad == conv over the big side), with the lrelu / relu / dropout
    derivative and the U-Net skip sum fused into the epilogue (gate / accumulate);
  * filter and bias gradients, losses and Adam are the kernels of csrc/train.cu;
  * parameters, gradients and Adam moments are single flat fp32 buffers ([G | D]), so one
    collective and one optimiser launch per network suffice.
"""
import ctypes as C

import torch

from advoc_b200 import _native as N
from advoc_b200 import nets
from advoc_b200.nets import _desc, _epilogue, _ptr, _stream


def _view_ptr(t, coff):
  return C.c_void_p(t.data_ptr() + 4 * coff)


class FlatParams(object):
  """All parameters in one flat buffer, generator first; the dict holds views."""

  def __init__(self, params, gen_prefix='generator', dis_prefix='discriminator'):
    names = sorted(params, key=lambda n: (0 if n.startswith(gen_prefix) else 1, n))
    self.names = names
    self.offsets, off = {}, 0
    for n in names:
      self.offsets[n] = off
      off += (params[n].numel() + 3) // 4 * 4   # keep every tensor 16-byte aligned
    self.total = off
    self.n_gen = min([self.offsets[n] for n in names if n.startswith(dis_prefix)] or [off])
    dev = params[names[0]].device
    self.p = torch.zeros(off, dtype=torch.float32, device=dev)
    self.g = torch.zeros(off, dtype=torch.float32, device=dev)
    self.m = torch.zeros(off, dtype=torch.float32, device=dev)
    self.v = torch.zeros(off, dtype=torch.float32, device=dev)
    self.P, self.G = {}, {}
    for n in names:
      o, k = self.offsets[n], params[n].numel()
      self.P[n] = self.p[o:o + k].view(params[n].shape)
      self.P[n].copy_(params[n])
      self.G[n] = self.g[o:o + k].view(params[n].shape)

  def range_of(self, pred):
    """[lo, hi) of the flat buffer covered by the (contiguous) tensors whose name satisfies pred."""
    sel = [n for n in self.names if pred(n)]
    if not sel:
      return 0, 0
    idx = [self.names.index(n) for n in sel]
    assert idx == list(range(idx[0], idx[0] + len(idx))), 'not contiguous in the flat buffer'
    last = sel[-1]
    return self.offsets[sel[0]], self.offsets[last] + (self.P[last].numel() + 3) // 4 * 4

  def gen_range(self):
    return 0, self.n_gen

  def dis_range(self):
    return self.n_gen, self.total


class TrainEngine(object):
  """One replica of the AdVoc train step for a fixed per-GPU batch."""

  def __init__(self, spec, ndf, params, batch, gan_weight=1.0, l1_weight=10.0, math=N.MATH_AUTO,
               lr=2e-4, beta1=0.5, beta2=0.999, eps=1e-8, process_group=None, world_size=1, rank=0,
               base_seed=0, overlap=True, use_graphs=True):
    """`rank` / `base_seed` key the dropout masks (every replica and every step draws its own);
    `overlap`: under data parallelism the gradient all-reduces run asynchronously behind the compute
    that does not depend on them (see train_loop).  `use_graphs`: with dropout='rng' or None the
    device work of d_step / g_step between two collectives is captured once and replayed as CUDA
    graphs (130-170 launches per step otherwise pace the step from the host on the small model);
    the Adam step size and the dropout counter live in device memory so the graphs follow them."""
    self.spec, self.ndf, self.B = spec, ndf, batch
    self.gan_weight, self.l1_weight = gan_weight, l1_weight
    self.lr, self.b1, self.b2, self.eps = lr, beta1, beta2, eps
    self.pg, self.world, self.rank = process_group, world_size, rank
    self.overlap = overlap
    self.use_graphs, self._graphs, self._graph_launches = use_graphs, {}, {}
    self.capture_launches = self.replayed_launches = 0    # launch accounting for bench.py
    # splitmix64 of (base_seed, rank): the per-replica dropout stream; step k uses seed0 + k
    z = (base_seed * 0x100000001B3 + rank + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    self.seed0 = 0 if (base_seed == 0 and rank == 0) else ((z ^ (z >> 31)) & 0xFFFFFFFF)
    self._pending_d = None
    self.collective_events = None   # bench.py: list of (start, end) CUDA events around every exposed wait
    self.flat = FlatParams(params)
    P = self.flat.P
    self.P = P
    dev = self.flat.p.device
    self.G = nets.Generator(spec, P, batch, math)
    self.Dr = nets.Discriminator(ndf, P, batch, math, spec.H[0], spec.W[0])
    self.Df = nets.Discriminator(ndf, P, batch, math, spec.H[0], spec.W[0])
    f32 = dict(dtype=torch.float32, device=dev)
    T, F = spec.H[0], spec.W[0]
    self.cat_real = torch.zeros((batch, T, F, 2), **f32)   # [x, target]
    self.cat_fake = torch.zeros((batch, T, F, 2), **f32)   # [x, generated]
    self.g_out = torch.zeros((batch, T, F, 1), **f32)      # d loss / d generated
    self.target_buf = torch.zeros((batch, T, F, 1), **f32)  # the L1 target (persistent: captured graphs read it)
    self.gCat = {k: torch.zeros_like(t) for k, t in self.G.Cat.items()}
    self.dz = [torch.zeros_like(t) for t in self.Dr.act]
    self.losses = torch.zeros(4, **f32)                    # d_loss, g_gan, g_l1, unused
    self.dz_fake = torch.zeros_like(self.dz[4])
    self.lr_t = torch.zeros(2, **f32)      # bias-corrected Adam step sizes (D, G) for advoc_adam_tf_step_dev
    self.seed_d = torch.zeros(1, dtype=torch.int64, device=dev)   # dropout step counter (advoc_epilogue.d_seed)
    self.t_d, self.t_g = 0, 0
    self.rnd = 0 if math == N.MATH_FP32 else 1   # TF32-round stored gradients for tensor-core consumers
    self.step_count = 0
    self._build_backward_geometry(math)
    self._build_buckets()
    self.refresh_weights()

  # -------------------------------------------------------------------------------------------
  def _build_backward_geometry(self, math):
    s, B, n = self.spec, self.B, self.spec.n_enc
    # conv geometry whose transposed form each decoder is, over the REAL (cropped) big side
    self.dec_b = {}
    for j, k in enumerate(range(n, 0, -1)):
      sh = 1 if j < s.n_stride1 else 2
      cin_small = self.G.Dk[k] + s.enc_ch[k - 1]
      ho, pt, _ = nets.same_pads(s.H[k - 1], 4, sh)
      wo, pl, _ = nets.same_pads(s.W[k - 1], 4, 2)
      assert ho == s.H[k] and wo == s.W[k]
      self.dec_b[k] = nets._Conv('generator/decoder_%d/conv2d_transpose' % k, 'conv',
                                 _desc(B, s.H[k - 1], s.W[k - 1], s.dec_ch[k], cin_small, sh, 2, pt, pl,
                                       ho, wo, math))
    # encoder / discriminator dgrad = transposed conv with the layer's own desc
    self.enc_t = {i: nets._Conv(self.G.enc[i].name, 'deconv', self.G.enc[i].desc)
                  for i in range(2, n + 1)}
    self.dis_t = [nets._Conv(L.name, 'deconv', L.desc) for L in self.Dr.layers]
    d0 = self.Dr.layers[0].desc
    # d(generated) from discriminator layer_1: transposed conv to ONE channel (channel 1 of the
    # 2-channel input), accumulated onto the L1 gradient
    self.dis_t0 = nets._Conv('discriminator/layer_1/conv2d', 'deconv',
                             _desc(B, d0.H, d0.W, 1, d0.Cout, d0.sh, d0.sw, d0.pad_t, d0.pad_l, d0.Ho,
                                   d0.Wo, N.MATH_FP32))

  def refresh_weights(self, which='GD'):
    """Re-derive the packed / TF32-rounded filter copies from the flat parameters after an Adam
    step; `which` names the network(s) whose parameters changed ('G', 'D' or 'GD')."""
    P = self.P
    if not hasattr(self, 'Wb'):
      self.Wb = {}
    # every derived copy is refreshed IN PLACE (captured graphs keep pointing at it); the forward and
    # the backward copy of a layer have different layouts, hence two tables
    if 'G' in which:
      self.G.prepare()
      for k, L in self.dec_b.items():      # deconv dgrad: conv over the big side
        self.Wb[L.name] = nets._pack_for_tc(L, P[L.name + '/kernel'], L.desc.Cin, self.Wb.get(L.name))
      for i, L in self.enc_t.items():      # conv dgrad: HWIO is already K-major, only round
        self.Wb[L.name] = nets._pack_for_tc(L, P[L.name + '/kernel'], L.desc.Cout, self.Wb.get(L.name))
    if 'D' in which:
      self.Dr.prepare()
      self.Df.Wp, self.Df.round = self.Dr.Wp, self.Dr.round
      for L in self.dis_t[1:]:
        self.Wb[L.name] = nets._pack_for_tc(L, P[L.name + '/kernel'], L.desc.Cout, self.Wb.get(L.name))
      k0 = P['discriminator/layer_1/conv2d/kernel']
      if not hasattr(self, 'w_sel'):
        self.w_sel = torch.empty_like(k0[:, :, 1:2, :].contiguous())
      self.w_sel.copy_(k0[:, :, 1:2, :])   # [4,4,1,ndf]: filter slice of input channel 1 (fixed address)

  def _wb(self, L):
    w = self.Wb.get(L.name)
    return w if w is not None else self.P[L.name + '/kernel']

  # -------------------------------------------------------------------------------------------
  # primitives
  # -------------------------------------------------------------------------------------------
  def _wgrad(self, desc, big, ld_big, coff_big, small, ld_small, coff_small, name):
    N.call('advoc_conv2d_wgrad', C.byref(desc), _view_ptr(big, coff_big), ld_big,
           _view_ptr(small, coff_small), ld_small, _ptr(self.flat.G[name + '/kernel']), _stream())

  def _bgrad(self, dy, ld, coff, pixels, channels, name):
    N.call('advoc_bias_grad', _view_ptr(dy, coff), ld, pixels, channels,
           _ptr(self.flat.G[name + '/bias']), _stream())

  def _run(self, L, x, ldx, coff_x, w, ep):
    fn = 'advoc_conv2d_fwd' if L.kind == 'conv' else 'advoc_conv2d_transpose_fwd'
    N.call(fn, C.byref(L.desc), _view_ptr(x, coff_x), ldx, _ptr(w), C.byref(ep), _stream())

  # -------------------------------------------------------------------------------------------
  # discriminator backward
  # -------------------------------------------------------------------------------------------
  def _d_backward(self, De, in_cat, param_grads, input_grad):
    """dz[4] holds d loss / d logits of layer_5.  Walks layers 5..1."""
    ch = De.chans
    for i in range(4, -1, -1):
      L = De.layers[i]
      d = L.desc
      dy = self.dz[i]
      x_in, ld_in = (in_cat, 2) if i == 0 else (De.act[i - 1], ch[i])
      if param_grads:
        self._wgrad(d, x_in, ld_in, 0, dy, ch[i + 1], 0, L.name)
        self._bgrad(dy, ch[i + 1], 0, d.N * d.Ho * d.Wo, ch[i + 1], L.name)
      if i > 0:
        Lt = self.dis_t[i]
        ep = _epilogue(None, self.dz[i - 1], ch[i], 0, N.ACT_NONE, gate=De.act[i - 1], ld_gate=ch[i],
                       gate_act=N.ACT_LRELU, round_tf32=self.rnd)
        self._run(Lt, dy, ch[i + 1], 0, self._wb(Lt), ep)
      elif input_grad:
        ep = _epilogue(None, self.g_out, 1, 0, N.ACT_NONE, accumulate=1)
        self._run(self.dis_t0, dy, ch[1], 0, self.w_sel, ep)

  # -------------------------------------------------------------------------------------------
  # generator backward (g_out holds d loss / d generated)
  # -------------------------------------------------------------------------------------------
  def _g_backward(self, x_cat, piece='all'):
    """piece: 'all', or one of the three stretches between which a gradient bucket is complete and its
    all-reduce can start (dec_8..5 + enc_8..5 hold 46 M of the regular model's 54 M
    parameters): 'dec' (decoder_1..n), 'enc_hi' (encoder_n..split), 'enc_lo' (encoder_(split-1)..1)."""
    s, n, G = self.spec, self.spec.n_enc, self.G
    keep = 0.5
    for k in (range(1, n + 1) if piece in ('all', 'dec') else ()):
      L = self.dec_b[k]
      d = L.desc
      if k == 1:
        dy, ld, co = self.g_out, 1, 0
      else:
        dy, ld, co = self.gCat[k - 1], self.gCat[k - 1].shape[3], 0
      cat = G.Cat[k]
      cc = cat.shape[3]
      self._wgrad(d, dy, ld, co, cat, cc, 0, L.name)
      self._bgrad(dy, ld, co, d.N * d.H * d.W, d.Cin, L.name)
      scale0 = (1.0 / keep) if (k < n and (k + 1) in s.dropout_decoders and self._dropout_on) else 1.0
      ep = _epilogue(None, self.gCat[k], cc, 0, N.ACT_NONE, gate=cat, ld_gate=cc, gate_act=N.ACT_RELU,
                     gate_split=G.Dk[k], gate_scale0=scale0, round_tf32=self.rnd)
      self._run(L, dy, ld, co, self._wb(L), ep)
    split = self._bucket_split
    enc_range = {'all': range(n, 0, -1), 'dec': (), 'enc_hi': range(n, split - 1, -1),
                 'enc_lo': range(split - 1, 0, -1)}[piece]
    for i in enc_range:
      L = G.enc[i]
      d = L.desc
      dy, ld, co = self.gCat[i], self.gCat[i].shape[3], G.Dk[i]
      if i == 1:
        x_in, ld_in = x_cat, 2
      else:
        x_in, ld_in = G.E[i - 1], s.enc_ch[i - 2]
      self._wgrad(d, x_in, ld_in, 0, dy, ld, co, L.name)
      self._bgrad(dy, ld, co, d.N * d.Ho * d.Wo, d.Cout, L.name)
      if i > 1:
        Lt = self.enc_t[i]
        prev = self.gCat[i - 1]
        ep = _epilogue(None, prev, prev.shape[3], G.Dk[i - 1], N.ACT_NONE, accumulate=1,
                       gate=G.E[i - 1], ld_gate=s.enc_ch[i - 2], gate_act=N.ACT_LRELU, round_tf32=self.rnd)
        self._run(Lt, dy, ld, co, self._wb(Lt), ep)

  # -------------------------------------------------------------------------------------------
  # optimiser + collective
  # -------------------------------------------------------------------------------------------
  def _allreduce(self, lo, hi):
    from advoc_b200 import dist as D
    D.allreduce_sum_(self.flat.g, lo, hi, self.pg, self.world)

  def _allreduce_async(self, lo, hi):
    """Starts the sum-all-reduce of flat.g[lo:hi] behind the work already queued on the current stream
    (NCCL runs it on its own stream) and returns a handle for _wait; None without peers."""
    from advoc_b200 import dist as D
    return D.allreduce_sum_async(self.flat.g, lo, hi, self.pg, self.world)

  def _wait(self, handles):
    """The current stream waits for the collectives; the time it idles is the EXPOSED all-reduce
    time (bench.py reads it from the recorded events)."""
    handles = [h for h in handles if h is not None]
    if not handles:
      return
    ev = None
    if self.collective_events is not None:
      ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
      ev[0].record()
    for h in handles:
      h.wait()
    if ev is not None:
      ev[1].record()
      self.collective_events.append(ev)

  def _bucketed(self):
    return self.world > 1 and self.overlap

  def _build_buckets(self):
    """Generator gradient buckets in the order the backward pass completes them: all decoders, then
    encoder_n..split, then encoder_(split-1)..1.  The flat buffer is sorted by name, so each is one
    contiguous range."""
    n = self.spec.n_enc
    self._bucket_split = 5 if n >= 5 else n
    f = self.flat
    enc_idx = lambda name: int(name.split('/')[1].split('_')[1])
    self._bucket = {
        'dec': f.range_of(lambda m: m.startswith('generator/decoder_')),
        'enc_hi': f.range_of(lambda m: m.startswith('generator/encoder_') and enc_idx(m) >= self._bucket_split),
        'enc_lo': f.range_of(lambda m: m.startswith('generator/encoder_') and enc_idx(m) < self._bucket_split),
    }
    lo, hi = f.gen_range()
    r = sorted(v for v in self._bucket.values() if v[1] > v[0])
    assert r[0][0] == lo and r[-1][1] == hi and all(a[1] == b[0] for a, b in zip(r, r[1:])), r

  def _adam(self, lo, hi, t):
    f = self.flat
    N.call('advoc_adam_tf_step', _view_ptr(f.p, lo), _view_ptr(f.g, lo), _view_ptr(f.m, lo),
           _view_ptr(f.v, lo), hi - lo, self.lr, self.b1, self.b2, self.eps, t, 1.0 / self.world,
           _stream())

  # -------------------------------------------------------------------------------------------
  # the two halves of `train_loop`
  # -------------------------------------------------------------------------------------------
  def restore_adam(self, m, v, t_d, t_g=None):
    """Resume the optimisers: `m` / `v` = {variable name: tensor} first / second moments (the
    `/Adam`, `/Adam_1` slots of a TF training checkpoint, advoc_b200.checkpoint.load_adam_slots),
    `t_d` / `t_g` = the number of D / G updates already applied (TF keeps them as beta powers;
    the reference bumps global_step in the G op only, advoc_model.py:253-255, and runs one D and one
    G update per loop, so both equal global_step)."""
    for n in self.flat.names:
      if n not in m or n not in v:
        raise KeyError('no Adam slots for %s' % n)
      o, k = self.flat.offsets[n], self.flat.P[n].numel()
      if tuple(m[n].shape) != tuple(self.flat.P[n].shape) or tuple(v[n].shape) != tuple(self.flat.P[n].shape):
        raise ValueError('Adam slot shape mismatch for %s' % n)
      self.flat.m[o:o + k].copy_(m[n].reshape(-1))
      self.flat.v[o:o + k].copy_(v[n].reshape(-1))
    self.t_d = int(t_d)
    self.t_g = int(t_d if t_g is None else t_g)
    self.step_count = self.t_d + self.t_g    # the dropout stream continues where the run left off

  def load_batch(self, x, target):
    """x, target f32 [B,T,513,1] on the device -> the discriminator input buffers."""
    self.cat_real[..., 0:1].copy_(x)
    self.cat_real[..., 1:2].copy_(target)
    self.cat_fake[..., 0:1].copy_(x)
    self.target_buf.copy_(target)

  # -------------------------------------------------------------------------------------------
  # CUDA-graph replay of the device work between two collectives
  # -------------------------------------------------------------------------------------------
  def _replay(self, key, fn, graphable=True):
    """Run `fn` (device work on persistent buffers only): eagerly the first time, then captured once
    and replayed.  Host-side state (step counters, input staging, collectives) stays outside `fn`."""
    if not (self.use_graphs and graphable):
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

  def _generate(self, dropout):
    """Generator forward into channel 1 of cat_fake.  dropout='rng' reads the step counter from device
    memory (self.seed_d, set by the caller before the launch / replay)."""
    self._dropout_on = dropout is not None
    # encoder_1 reads x as channel 0 of the 2-channel buffer; decoder_1 writes channel 1
    return self.G.forward(self.cat_fake, out=self.cat_fake, out_ld=2, out_coff=1, dropout=dropout,
                          seed_dev=self.seed_d if dropout == 'rng' else None, x_ld=2)

  def _bump_seed(self):
    self.step_count += 1
    self.seed_d.fill_(self.seed0 + self.step_count)

  def _finish_d(self):
    """Second half of a deferred D update: wait for its all-reduce, Adam, refresh the packed filters."""
    if self._pending_d is None:
      return
    handles, lo, hi = self._pending_d
    self._pending_d = None
    self._wait(handles)
    self._apply('D', lo, hi)

  def _apply(self, which, lo, hi):
    """TF1-Adam on one network's slice + in-place refresh of its packed filter copies (one graph)."""
    if which == 'D':
      self.t_d += 1
      t, slot = self.t_d, 0
    else:
      self.t_g += 1
      t, slot = self.t_g, 1
    self.lr_t[slot:slot + 1].fill_(self.lr * (1.0 - self.b2 ** t) ** 0.5 / (1.0 - self.b1 ** t))

    def update():
      f = self.flat
      N.call('advoc_adam_tf_step_dev', _view_ptr(f.p, lo), _view_ptr(f.g, lo), _view_ptr(f.m, lo),
             _view_ptr(f.v, lo), hi - lo, _view_ptr(self.lr_t, slot), self.b1, self.b2, self.eps,
             1.0 / self.world, _stream())
      self.refresh_weights(which)
    self._replay('adam_' + which, update)

  def _d_compute(self, dropout):
    lo, hi = self.flat.dis_range()
    self.flat.g[lo:hi].zero_()
    self.losses[0:1].zero_()
    self._generate(dropout)
    p_real = self.Dr.forward(self.cat_real)
    p_fake = self.Df.forward(self.cat_fake)
    N.call('advoc_gan_logloss', _ptr(p_real), _ptr(p_fake), p_real.numel(), 0, 1.0, _ptr(self.losses),
           _ptr(self.dz[4]), _ptr(self.dz_fake), _stream())
    self._d_backward(self.Dr, self.cat_real, True, False)
    self.dz[4].copy_(self.dz_fake)
    self._d_backward(self.Df, self.cat_fake, True, False)

  def d_step(self, x, target, dropout='rng', apply=True, defer=False):
    """advoc_model.py:257 `D_train_op` on one minibatch.  defer: leave the all-reduce in flight and the
    optimiser update pending (train_loop finishes it under the next generator forward)."""
    self._finish_d()
    lo, hi = self.flat.dis_range()
    self.load_batch(x, target)
    self._bump_seed()
    graphable = dropout == 'rng' or dropout is None
    self._replay(('d', dropout if graphable else None), lambda: self._d_compute(dropout), graphable)
    if apply and defer and self._bucketed():
      self._pending_d = ([self._allreduce_async(lo, hi)], lo, hi)
      return
    self._wait([self._allreduce_async(lo, hi)])
    if apply:
      self._apply('D', lo, hi)

  def _g_forward(self, dropout):
    lo, hi = self.flat.gen_range()
    self.flat.g[lo:hi].zero_()
    self.losses[1:3].zero_()
    self._generate(dropout)
    N.call('advoc_l1_loss', _ptr(self.cat_fake), 2, 1, _ptr(self.target_buf), self.g_out.numel(), self.l1_weight,
           _view_ptr(self.losses, 2), _ptr(self.g_out), 0, _stream())

  def _g_through_d(self):
    p_fake = self.Df.forward(self.cat_fake)
    N.call('advoc_gan_logloss', None, _ptr(p_fake), p_fake.numel(), 1, self.gan_weight,
           _view_ptr(self.losses, 1), None, _ptr(self.dz[4]), _stream())
    self._d_backward(self.Df, self.cat_fake, False, True)

  def g_step(self, x, target, dropout='rng', apply=True):
    """advoc_model.py:254-255 `G_train_op` on one minibatch (bumps the global step)."""
    lo, hi = self.flat.gen_range()
    self.load_batch(x, target)
    self._bump_seed()
    graphable = dropout == 'rng' or dropout is None
    mode = dropout if graphable else None
    self._replay(('g_fwd', mode), lambda: self._g_forward(dropout), graphable)
    # a deferred D update (its all-reduce ran under the generator forward above) lands here: the
    # discriminator pass below goes through the already-updated D like the reference's (advoc_model.py:285-289)
    self._finish_d()
    if self._bucketed():
      # the backward pass is cut where a gradient bucket completes; the all-reduces between the pieces stay eager
      if self.gan_weight > 0:
        self._replay(('g_dis', mode), self._g_through_d, graphable)
      handles = []
      for piece in ('dec', 'enc_hi', 'enc_lo'):
        self._replay(('g_bwd_' + piece, mode), lambda piece=piece: self._g_backward(self.cat_fake, piece), graphable)
        handles.append(self._allreduce_async(*self._bucket[piece]))
    else:
      def rest():
        if self.gan_weight > 0:
          self._g_through_d()
        self._g_backward(self.cat_fake, 'all')
      self._replay(('g_bwd', mode), rest, graphable)
      handles = [self._allreduce_async(lo, hi)]
    self._wait(handles)
    if apply:
      self._apply('G', lo, hi)
    return self.t_g

  def train_loop(self, batch_d, batch_g, dropout='rng'):
    """One reference `train_loop` (advoc_model.py:285-289): D step (skipped when gan_weight <= 0)
    then G step, each on its own minibatch (x, target)."""
    if self.gan_weight > 0:
      self.d_step(batch_d[0], batch_d[1], dropout, defer=True)
    return self.g_step(batch_g[0], batch_g[1], dropout)

  def loss_values(self):
    """(d_loss, g_loss_GAN * gan_weight, g_loss_L1 * l1_weight) of the last steps (syncs)."""
    self._finish_d()
    v = self.losses.tolist()
    N.raise_if_aborted('TrainEngine')
    return v[0], v[1], v[2]
