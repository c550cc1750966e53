"""Host-side mirror of the reference's model interface (models/advoc/{model,util,advoc_model,
advoc_model_small,spectral_util}.py) on top of the B200 engine in `advoc_b200.nets`.

Same class names, class-attribute hyper-parameters and method names as the reference; where the
reference builds a TF1 graph node, these methods run the sm_100a kernels eagerly on
`torch.Tensor` containers (NHWC [b, time, freq, ch], float32, CUDA).
"""
import numpy as np
import torch

from advoc_b200 import _native as N
from advoc_b200 import nets
from advoc_b200 import spectral

EPS = 1e-12  # models/advoc/advoc_model.py:8


class Modes(object):
  """models/advoc/model.py:15-18"""
  TRAIN = 'train'
  EVAL = 'eval'
  INFER = 'infer'


class Model(object):
  """models/advoc/model.py:1-12"""

  def __init__(self, mode, *args, **kwargs):
    self.mode = mode

  def __call__(self):
    raise Exception('Abstract method')

  def train_loop(self):
    raise Exception('Abstract method')

  def eval_ckpt(self, ckpt_fp):
    raise Exception('Abstract method')


def override_model_attrs(model, overrides):
  """`--model_overrides "k=v,k=v"`, each value cast to the type of the class attribute.
  reference: models/advoc/util.py:1-20."""
  if overrides is not None and len(overrides.strip()):
    for key, val in [p.split('=') for p in overrides.split(',')]:
      val_type = type(getattr(model, key))
      if val_type == bool:
        setattr(model, key, val in ['True', 'true', 't', '1'])
      elif val_type == list:
        setattr(model, key, val.split(';'))
      else:
        setattr(model, key, val_type(val))
  attrs = sorted(x for x in dir(model) if not x.startswith('_') and not callable(getattr(model, x)))
  summary = '\n'.join('{},{}'.format(k, getattr(model, k)) for k in attrs)
  return model, summary


class SpectralUtil(object):
  """models/advoc/spectral_util.py:6-60: mel <-> magnitude matmuls with fixed matrices."""
  NFFT = 1024
  NHOP = 256
  FMIN = 125.
  FMAX = 7600.
  NMELS = 80
  fs = 22050

  def __init__(self, n_mels=80, fs=22050, device=None):
    self.NMELS = n_mels
    self.fs = fs
    self.meltrans_np = spectral.create_mel_filterbank(
        self.fs, self.NFFT, fmin=self.FMIN, fmax=self.FMAX, n_mels=self.NMELS)
    self.invmeltrans_np = spectral.create_inverse_mel_filterbank(
        self.fs, self.NFFT, fmin=self.FMIN, fmax=self.FMAX, n_mels=self.NMELS)
    dev = device or torch.device('cuda', torch.cuda.current_device())
    self.meltrans = torch.from_numpy(self.meltrans_np.astype(np.float32)).to(dev)        # [80,513]
    self.invmeltrans = torch.from_numpy(self.invmeltrans_np.astype(np.float32)).to(dev)  # [513,80]

  def mag_to_mel_linear_spec(self, mag_spec):
    """[b,t,513,1] -> [b,t,80,1]  (spectral_util.py:29-32)"""
    return spectral.matmul_lastdim(mag_spec[:, :, :, 0], self.meltrans).unsqueeze(-1)

  def mel_linear_to_mag_spec(self, mel_spec, transform='inverse'):
    """[b,t,80,1] -> [b,t,513,1]; no >=0 clamp  (spectral_util.py:34-43)"""
    if transform == 'inverse':
      w = self.invmeltrans
    elif transform == 'transposed':
      w = self.meltrans.t().contiguous()
    else:
      raise NotImplementedError()
    return spectral.matmul_lastdim(mel_spec[:, :, :, 0], w).unsqueeze(-1)

  def tacotron_mel_to_mag(self, X_mel_dbnorm):
    """dB-normalised mel [T,80] -> magnitude [T,513]: 10^((x*100-100+20)/20) . pinv(W)^T fused
    in one kernel (spectral_util.py:52-60, scripts/spectrogram_advoc.py:15-22).  Accepts a numpy
    array (returns numpy f64 like the reference) or a CUDA tensor (returns a tensor)."""
    if isinstance(X_mel_dbnorm, np.ndarray):
      x = torch.from_numpy(np.ascontiguousarray(X_mel_dbnorm, dtype=np.float32)).to(
          self.invmeltrans.device)
      return spectral.matmul_lastdim(x, self.invmeltrans, pow10_scale=True).cpu().numpy().astype(
          np.float64)
    return spectral.matmul_lastdim(X_mel_dbnorm, self.invmeltrans, pow10_scale=True)


class Advoc(Model):
  """models/advoc/advoc_model.py:10-22 (hyper-parameters are class attributes)."""
  audio_fs = 22050
  subseq_len = 256
  n_mels = 80
  ngf = 64
  ndf = 64
  gan_weight = 1.
  l1_weight = 10.
  train_batch_size = 8
  eval_batch_size = 1
  separable_conv = False
  use_batchnorm = False
  generator_type = "pix2pix"
  # engine-side knobs (not in the reference)
  math_mode = N.MATH_AUTO
  _n_enc = 8
  _dropout_decoders = (8, 7, 6)

  def __init__(self, mode, params=None, seed=0):
    super(Advoc, self).__init__(mode)
    self.params = params
    self._seed = seed
    self._gen = {}
    self._dis = {}
    self._step_count = 0

  # -- parameters ---------------------------------------------------------
  def init_params(self, seed=None):
    self.params = nets.init_params(self.ngf, self.ndf, self._n_enc,
                                   self._seed if seed is None else seed)
    return self.params

  def gen_spec(self):
    return nets.GenSpec(self.ngf, self._n_enc, self._dropout_decoders, self.subseq_len)

  def _check_supported(self):
    if self.separable_conv or self.use_batchnorm:
      raise NotImplementedError('separable_conv / use_batchnorm variants are out of scope')
    if self.generator_type != 'pix2pix':
      raise NotImplementedError(self.generator_type)
    if self.params is None:
      self.init_params()

  def _generator(self, batch, math=None):
    math = self.math_mode if math is None else math
    if (batch, math) not in self._gen:
      self._gen[(batch, math)] = nets.Generator(self.gen_spec(), self.params, batch, math)
    return self._gen[(batch, math)]

  def _discriminator(self, batch):
    if batch not in self._dis:
      self._dis[batch] = nets.Discriminator(self.ndf, self.params, batch, self.math_mode,
                                            self.subseq_len)
    return self._dis[batch]

  # -- the reference's graph builders, executed eagerly --------------------
  def build_generator(self, x, dropout='rng', seed=None):
    """x [b, subseq_len, 513, 1] -> generated magnitude spectrogram, same shape.
    advoc_model.py:75-166.  Dropout on decoder_8/7/6 is active in every mode like the
    reference (:144-149); pass dropout=None to disable it or a {k: mask} dict to inject."""
    self._check_supported()
    g = self._generator(x.shape[0])
    self._step_count += 1
    return g.forward(x.contiguous(), dropout=dropout,
                     seed=self._step_count if seed is None else seed)

  def build_discriminator(self, discrim_inputs, discrim_targets):
    """advoc_model.py:168-204: sigmoid patch map [b,30,62,1]."""
    self._check_supported()
    d = self._discriminator(discrim_inputs.shape[0])
    return d.forward(torch.cat([discrim_inputs, discrim_targets], dim=3).contiguous())

  # -- training graph (advoc_model.py:206-289) ------------------------------------------
  def __call__(self, x, target=None, x_wav=None, x_mel_spec=None, process_group=None, world_size=1, rank=0):
    """The reference builds generator, both discriminator towers, the three losses, the variable
    partition and the two Adam optimisers here (advoc_model.py:206-257); this binds the same objects
    onto a `train.TrainEngine` for the batch size of `x`.

    x / target: tensors [b, subseq_len, 513, 1] (inverted-mel magnitude input, true magnitude) -- one
    fixed minibatch, re-used by every update -- or `x` = an iterator yielding (x, target) pairs, the
    eager stand-in for the reference's graph tensors fed by `decode_extract_and_batch`: every D / G
    update then pulls its own minibatch, like every `sess.run` of the reference.  x_wav / x_mel_spec only
    feed tf.summary nodes in the reference (:258-281) and are accepted and ignored.
    Under data parallelism pass the process group / world size / rank (torchrun: advoc_b200.dist.init)."""
    from advoc_b200.train import TrainEngine
    self._check_supported()
    self.spectral = SpectralUtil(n_mels=self.n_mels, fs=self.audio_fs)
    if target is None:
      self._batches = iter(x)
      first = next(self._batches)
      self._pushback = [first]
      batch = int(first[0].shape[0])
    else:
      self._batches, self._pushback = None, []
      self._fixed = (x, target)
      batch = int(x.shape[0])
    self._engine = TrainEngine(self.gen_spec(), self.ndf, self.params, batch, gan_weight=self.gan_weight,
                               l1_weight=self.l1_weight, math=self.math_mode, process_group=process_group,
                               world_size=world_size, rank=rank, base_seed=self._seed)
    self.params = self._engine.P          # views into the engine's flat parameter buffer
    self._gen, self._dis = {}, {}
    self.G_vars = [n for n in self._engine.flat.names if n.startswith('generator')]
    self.D_vars = [n for n in self._engine.flat.names if n.startswith('discriminator')]
    self.step = 0
    self.D_train_op = lambda: self._engine.d_step(*self._next_batch())
    self.G_train_op = lambda: self._engine.g_step(*self._next_batch())
    return self

  def _next_batch(self):
    if self._batches is None:
      return self._fixed
    if self._pushback:
      return self._pushback.pop()
    return next(self._batches)

  def train_loop(self, sess=None):
    """advoc_model.py:285-289: D update (skipped when gan_weight <= 0) on one minibatch, then G update
    on the next; returns the global step (bumped by the G op only, :253-255).  `sess` is the
    reference's session argument and is ignored."""
    if getattr(self, '_engine', None) is None:
      raise Exception('call the model on a batch first (advoc_model.py:206)')
    e = self._engine
    bd = self._next_batch() if self.gan_weight > 0 else None
    bg = self._next_batch()
    self.step = e.train_loop(bd, bg)
    return self.step

  def losses(self):
    """(disc_loss, gen_loss_GAN * gan_weight, gen_loss_L1 * l1_weight) of the last updates: the scalars
    the reference writes as tf.summary (advoc_model.py:271-274)."""
    return self._engine.loss_values()


class AdvocSmall(Advoc):
  """models/advoc/advoc_model_small.py:14-23,107-108,128-134."""
  ngf = 32
  ndf = 32
  num_enc_layers = 4
  _dropout_decoders = (5, 4)

  @property
  def _n_enc(self):
    return self.num_enc_layers + 1
