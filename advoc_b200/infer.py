"""Batched mel -> magnitude-spectrogram inference: the reference's chunk loop
(scripts/spectrogram_advoc.py:80-95: batch-1 `sess.run` per 256-frame chunk, host->device copy
per chunk) as ONE batched forward on persistent buffers, replayed from a CUDA graph.

Pipeline per call (all on the current stream):
  host mel [n, T, 80] --H2D--> (dB de-normalise) x pinv(mel_fb)^T  (one fused kernel: generator
  "layer 0", models/advoc/spectral_util.py:34-43 / scripts/spectrogram_advoc.py:15-22)
  --> U-Net generator --> magnitude [n, T, 513] --D2H--> pinned host buffer.
"""
import numpy as np
import torch

from advoc_b200 import _native as N
from advoc_b200 import nets
from advoc_b200 import spectral
from advoc_b200.model import SpectralUtil


class MelToMag(object):
  """Fixed-batch engine.  `input_kind`: 'linear' (linear-amplitude mel, the training-time
  input of models/advoc/train_evaluate.py:53-56) or 'dbnorm' (r9y9 dB-normalised mel in [0,1],
  the `.npy` inputs of scripts/spectrogram_advoc.py)."""

  def __init__(self, model, batch, input_kind='linear', dropout='rng', use_graph=True, math=N.MATH_F16):
    """math: MATH_F16 (default: fp16 operand storage between the tensor-core layers, same 10-bit
    mantissa as TF32 at half the bytes; within the 1e-3 parity budget, tests/test_gpu_nets.py),
    MATH_AUTO (TF32 operands, fp32 storage) or MATH_FP32 (CUDA cores)."""
    model._check_supported()
    self.model, self.B, self.kind, self.dropout = model, batch, input_kind, dropout
    self.T, self.n_mels = model.subseq_len, model.n_mels
    self.su = SpectralUtil(n_mels=model.n_mels, fs=model.audio_fs)
    self.G = model._generator(batch, math)
    self.G.prepare()
    dev = self.G.dev
    # dropout step counter in device memory: bumped INSIDE the captured graphs, read by the decoder
    # epilogues (advoc_epilogue.d_seed), so every replay draws fresh masks like every sess.run of the
    # reference does (advoc_model.py:144-149)
    self.seed_d = torch.zeros(1, dtype=torch.int64, device=dev)
    self.mel_d = torch.zeros((batch, self.T, self.n_mels), dtype=torch.float32, device=dev)
    self.x_d = torch.empty((batch, self.T, 513, 1), dtype=torch.float32, device=dev)
    self.mel_h = torch.zeros((batch, self.T, self.n_mels), dtype=torch.float32).pin_memory()
    self.out_h = torch.empty((batch, self.T, 513), dtype=torch.float32).pin_memory()
    self.use_graph = use_graph
    self._graph = None
    self._seed = 0
    self.launches_per_step = None

  # -- device-resident step (inputs already in HBM) ------------------------
  def _launch(self, seed, n_valid=None):
    """seed: host value for the eager path, None = bump and read the device counter."""
    spectral_ptr = spectral._ptr
    N.call('advoc_matmul_lastdim_f32', spectral_ptr(self.mel_d), spectral_ptr(self.su.invmeltrans),
           spectral_ptr(self.x_d), self.B * self.T, self.n_mels, 513,
           1 if self.kind == 'dbnorm' else 0, -100.0, 20.0, spectral._stream())
    if n_valid is not None and n_valid < self.B * self.T:
      # the reference zero-pads in the magnitude domain, after the pinv lift
      # (scripts/spectrogram_advoc.py:81-84)
      self.x_d.view(self.B * self.T, 513)[n_valid:].zero_()
    if seed is None:
      self.seed_d.add_(1)
      return self.G.forward(self.x_d, dropout=self.dropout, seed_dev=self.seed_d)
    return self.G.forward(self.x_d, dropout=self.dropout, seed=seed)

  def step_device(self, n_valid=None):
    """One forward over the batch resident in `self.mel_d`; returns the device output
    [B, T, 513, 1].  Replays a CUDA graph after the first (capturing) call.  Every call -- eager or
    replayed -- advances the dropout step counter, so call k of an engine uses the masks of seed k on
    either path (tests/test_gpu_pipeline.py)."""
    self._seed += 1
    if not self.use_graph or n_valid is not None:
      self.seed_d.fill_(self._seed)     # keep the device counter in step with the host one
      return self._launch(self._seed, n_valid)
    if self._graph is None:
      n0 = N.launch_count()
      self._launch(self._seed)  # warm-up outside capture (lazy module load, caches)
      self.launches_per_step = N.launch_count() - n0
      torch.cuda.synchronize()
      g = torch.cuda.CUDAGraph()
      with torch.cuda.graph(g):
        self._out = self._launch(None)
      self._graph = g
      self.seed_d.fill_(self._seed - 1)   # the replay below bumps it to this call's seed
    self._graph.replay()
    return self._out

  # -- streaming host API: copies overlap the forward of the neighbouring batches ------------
  def _stream_setup(self):
    """Two input / output buffer sets, one captured graph per set, one stream per copy direction
    (a shared copy stream would queue the next batch's input behind the previous batch's result and
    with it serialise compute and device->host transfer) and events."""
    if getattr(self, '_sets', None) is not None:
      return
    dev = self.G.dev
    self._h2d_stream = torch.cuda.Stream(device=dev)
    self._d2h_stream = torch.cuda.Stream(device=dev)
    self._sets = []
    for k in range(2):
      s = dict(mel_d=torch.zeros_like(self.mel_d), out_d=torch.empty((self.B, self.T, 513, 1), dtype=torch.float32,
                                                                     device=dev),
               mel_h=torch.zeros((self.B, self.T, self.n_mels), dtype=torch.float32).pin_memory(),
               out_h=torch.empty((self.B, self.T, 513), dtype=torch.float32).pin_memory(),
               h2d=torch.cuda.Event(), done=torch.cuda.Event(), d2h=torch.cuda.Event(), graph=None)
      self._sets.append(s)

  def _launch_set(self, s, seed):
    spectral_ptr = spectral._ptr
    N.call('advoc_matmul_lastdim_f32', spectral_ptr(s['mel_d']), spectral_ptr(self.su.invmeltrans),
           spectral_ptr(self.x_d), self.B * self.T, self.n_mels, 513,
           1 if self.kind == 'dbnorm' else 0, -100.0, 20.0, spectral._stream())
    if seed is None:
      self.seed_d.add_(1)
      return self.G.forward(self.x_d, out=s['out_d'], dropout=self.dropout, seed_dev=self.seed_d)
    return self.G.forward(self.x_d, out=s['out_d'], dropout=self.dropout, seed=seed)

  def run_stream(self, batches):
    """Generator over host batches [B, T, n_mels] (float32; pinned or not): yields one pinned host
    tensor [B, T, 513] per input batch, in order, valid until the generator is advanced again.
    The host->device copy of batch i+1 and the device->host copy of batch i-1 run on their own
    streams while batch i is in the generator (the reference's chunk loop, scripts/spectrogram_advoc.py:88-92,
    pays both copies serially for every 256-frame chunk)."""
    self._stream_setup()
    main = torch.cuda.current_stream()
    hs, ds = self._h2d_stream, self._d2h_stream
    pending = []
    i = -1
    for i, mel in enumerate(batches):
      s = self._sets[i & 1]
      if isinstance(mel, np.ndarray):
        mel = torch.from_numpy(mel)
      mel = mel.reshape(self.B, self.T, self.n_mels)
      if mel.dtype != torch.float32:
        raise ValueError()
      if i >= 2:
        s['d2h'].synchronize()          # the result that used this buffer set has reached the host
        N.raise_if_aborted('MelToMag.run_stream')
        yield pending.pop(0)
      if not mel.is_pinned():
        s['mel_h'].copy_(mel)
        mel = s['mel_h']
      # buffer set reuse is safe: the host waited above for the result of batch i-2, i.e. for everything
      # that read this set's input or wrote its output
      with torch.cuda.stream(hs):
        s['mel_d'].copy_(mel, non_blocking=True)
        s['h2d'].record(hs)
      main.wait_event(s['h2d'])
      self._seed += 1
      if self.use_graph:
        if s['graph'] is None:
          self._launch_set(s, self._seed)        # warm-up outside capture
          torch.cuda.synchronize()
          g = torch.cuda.CUDAGraph()
          with torch.cuda.graph(g):
            self._launch_set(s, None)
          s['graph'] = g
          self.seed_d.fill_(self._seed - 1)
        s['graph'].replay()
      else:
        self.seed_d.fill_(self._seed)
        self._launch_set(s, self._seed)
      s['done'].record(main)
      with torch.cuda.stream(ds):
        ds.wait_event(s['done'])
        s['out_h'].copy_(s['out_d'].view(self.B, self.T, 513), non_blocking=True)
        s['d2h'].record(ds)
      pending.append(s['out_h'])
    for k, out in enumerate(pending):
      self._sets[(i - len(pending) + 1 + k) & 1]['d2h'].synchronize()
      N.raise_if_aborted('MelToMag.run_stream')
      yield out

  # -- public host API ----------------------------------------------------
  def __call__(self, mel, n_valid=None):
    """mel: host float32 array/tensor [B, T, n_mels] (or [B, T, n_mels, 1]); frames at flat
    index >= n_valid are treated as the reference's zero padding.  Returns the pinned host
    tensor [B, T, 513] (valid until the next call)."""
    if isinstance(mel, np.ndarray):
      mel = torch.from_numpy(mel)
    mel = mel.reshape(self.B, self.T, self.n_mels)
    if mel.dtype != torch.float32:
      raise ValueError()
    if not mel.is_pinned():
      self.mel_h.copy_(mel)
      mel = self.mel_h
    self.mel_d.copy_(mel, non_blocking=True)
    out = self.step_device(n_valid)
    self.out_h.copy_(out.view(self.B, self.T, 513), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    N.raise_if_aborted('MelToMag.__call__')
    return self.out_h


def mel_to_mag(model, mel, input_kind='dbnorm', batch=None, dropout='rng'):
  """Whole-utterance helper: mel [T_total, 80] (numpy) -> magnitude [T_total, 513] (numpy f32),
  following the reference's pad rule (always pads to floor(T/256)*256 + 256 frames,
  scripts/spectrogram_advoc.py:82-86) and trim (:94)."""
  T = model.subseq_len
  n_frames = mel.shape[0]
  padded = n_frames - n_frames % T + T
  buf = np.zeros((padded, mel.shape[1]), dtype=np.float32)
  buf[:n_frames] = mel
  chunks = buf.reshape(padded // T, T, mel.shape[1])
  n = chunks.shape[0]
  eng = MelToMag(model, n if batch is None else batch, input_kind, dropout, use_graph=False)
  out = np.empty((n, T, 513), dtype=np.float32)
  for i in range(0, n, eng.B):
    blk = chunks[i:i + eng.B]
    if blk.shape[0] < eng.B:
      blk = np.concatenate([blk, np.zeros((eng.B - blk.shape[0],) + blk.shape[1:], np.float32)])
    valid = max(0, min(eng.B * T, n_frames - i * T))
    out[i:i + eng.B] = eng(blk, n_valid=valid).numpy()[:min(eng.B, n - i)]
  return out.reshape(padded, 513)[:n_frames]
