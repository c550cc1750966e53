"""Drop-in for the reference's `advoc.spectral` feature path on B200.

Same function names, argument meaning, shapes, dtypes and exceptions as
/root/reference/advoc/spectral.py; the arithmetic runs in the sm_100a kernels of
libadvoc_b200.so (csrc/spectral.cu).  numpy entry points take/return numpy arrays (host
buffers, copies included); the `*_tf` twins of the reference's TensorFlow graph builders take
and return `torch.Tensor`s on the GPU (container type only) in the reference's
[b, n, feats, ch] convention.

There is no CPU fallback: without the native library or a GPU these functions raise.
"""
import ctypes as C
from functools import lru_cache

import numpy as np
import torch

from advoc_b200 import _native as N


def _stream():
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
  return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _device():
  if not torch.cuda.is_available():
    raise RuntimeError('advoc_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
  return torch.device('cuda', torch.cuda.current_device())


# ---------------------------------------------------------------------------
# host-side constants (window, twiddles, mel filterbank): computed once in f64, cached
# ---------------------------------------------------------------------------
def lws_hann_default(nfft, nhop, dtype=np.float32):
  """Analysis window of `lws.lws(nfft, nhop)`: sqrt(hann_offset * 2 * hop / nfft).

  reference: advoc/spectral.py:44-57.  Returns a numpy array (the reference returns a TF
  constant); `dtype` may be a numpy or torch dtype.
  """
  i = np.arange(nfft, dtype=np.float64) + 0.5
  w = np.sqrt((0.5 - 0.5 * np.cos(2.0 * np.pi * i / nfft)) * (2.0 * nhop / nfft))
  if isinstance(dtype, torch.dtype):
    return torch.from_numpy(w).to(dtype)
  return w.astype(dtype)


def _slaney_hz(mels):
  mels = np.asarray(mels, dtype=np.float64)
  f_sp = 200.0 / 3.0
  brk_mel = 1000.0 / f_sp
  step = np.log(6.4) / 27.0
  return np.where(mels < brk_mel, mels * f_sp, 1000.0 * np.exp(step * (mels - brk_mel)))


def _slaney_mel(hz):
  hz = np.asarray(hz, dtype=np.float64)
  f_sp = 200.0 / 3.0
  step = np.log(6.4) / 27.0
  return np.where(hz < 1000.0, hz / f_sp,
                  1000.0 / f_sp + np.log(np.maximum(hz, 1e-30) / 1000.0) / step)


@lru_cache(maxsize=8)
def create_mel_filterbank(fs, nfft, fmin=0.0, fmax=None, n_mels=128):
  """Slaney-scale, area-normalised triangular filterbank, f64 [n_mels, nfft//2+1].

  reference: advoc/spectral.py:86-88 (`librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`,
  librosa 0.6.3 defaults htk=False, norm=1).
  """
  fmax = fs / 2.0 if fmax is None else fmax
  freqs = np.linspace(0.0, fs / 2.0, nfft // 2 + 1)
  pts = _slaney_hz(np.linspace(_slaney_mel(fmin), _slaney_mel(fmax), n_mels + 2))
  lower = (freqs[None, :] - pts[:-2, None]) / (pts[1:-1] - pts[:-2])[:, None]
  upper = (pts[2:, None] - freqs[None, :]) / (pts[2:] - pts[1:-1])[:, None]
  fb = np.maximum(0.0, np.minimum(lower, upper))
  fb *= (2.0 / (pts[2:] - pts[:-2]))[:, None]
  fb.setflags(write=False)
  return fb


@lru_cache(maxsize=8)
def create_inverse_mel_filterbank(fs, nfft, fmin=0.0, fmax=None, n_mels=128):
  """reference: advoc/spectral.py:91-94.  f64 [nfft//2+1, n_mels]."""
  inv = np.linalg.pinv(create_mel_filterbank(fs, nfft, fmin=fmin, fmax=fmax, n_mels=n_mels))
  inv.setflags(write=False)
  return inv


class _Consts(object):
  """Device-resident window / twiddle / filterbank tables, keyed by their parameters."""
  _cache = {}

  @classmethod
  def stft(cls, nfft, nhop, dev):
    key = ('stft', nfft, nhop, dev)
    if key not in cls._cache:
      win = torch.from_numpy(lws_hann_default(nfft, nhop, np.float32)).to(dev)
      j = np.arange(nfft, dtype=np.float64)
      tw = np.stack([np.cos(2 * np.pi * j / nfft), -np.sin(2 * np.pi * j / nfft)], axis=1)
      cls._cache[key] = (win, torch.from_numpy(tw.astype(np.float32)).to(dev))
    return cls._cache[key]

  @classmethod
  def mel(cls, fs, nfft, fmin, fmax, n_mels, dev):
    key = ('mel', fs, nfft, fmin, fmax, n_mels, dev)
    if key not in cls._cache:
      fb64 = create_mel_filterbank(fs, nfft, fmin=fmin, fmax=fmax, n_mels=n_mels)
      fb = torch.from_numpy(fb64.astype(np.float32)).to(dev).contiguous()
      ranges = torch.empty((n_mels, 2), dtype=torch.int32, device=dev)
      N.call('advoc_mel_ranges', _ptr(fb), n_mels, nfft // 2 + 1, _ptr(ranges), _stream())
      cls._cache[key] = (fb, ranges)
    return cls._cache[key]

  @classmethod
  def inv_mel(cls, fs, nfft, fmin, fmax, n_mels, dev):
    key = ('inv', fs, nfft, fmin, fmax, n_mels, dev)
    if key not in cls._cache:
      inv = create_inverse_mel_filterbank(fs, nfft, fmin=fmin, fmax=fmax, n_mels=n_mels)
      cls._cache[key] = torch.from_numpy(inv.astype(np.float32)).to(dev).contiguous()
    return cls._cache[key]


def num_frames(nsamps, nfft, nhop, pad_end=True, tf_rule=False):
  """Frame-count rule of advoc/spectral.py:32-39 (and lws' own, tests/test_spectral.py:35-36);
  tf_rule: without pad_end, tf.contrib.signal.stft's floor((n - nfft) / hop) + 1 whole frames."""
  return N.lib().advoc_num_frames(int(nsamps), int(nfft), int(nhop), 1 if pad_end else (2 if tf_rule else 0))


# ---------------------------------------------------------------------------
# device-level ops on torch tensors
# ---------------------------------------------------------------------------
def stft_tf(x, nfft, nhop, pad_end=True, _lws_rule=False):
  """Batched STFT.  x f32 [b, nsamps, 1, nch] (cuda) -> c64 [b, frames, nfft//2+1, nch].

  reference: advoc/spectral.py:60-83.  (_lws_rule: the numpy entry point `stft` frames like lws when
  pad_end is off -- the last partial frame zero-padded -- instead of like tf.contrib.signal.stft.)
  """
  if x.dim() != 4 or x.shape[2] != 1:
    raise ValueError()
  if x.dtype != torch.float32:
    raise ValueError()
  x = x.contiguous()
  b, nsamps, _, nch = x.shape
  win, tw = _Consts.stft(nfft, nhop, x.device)
  # without pad_end tf.contrib.signal.stft keeps whole frames only: floor((n - nfft) / hop) + 1, possibly 0
  frames = num_frames(nsamps, nfft, nhop, pad_end, tf_rule=not _lws_rule)
  out = torch.empty((b, frames, nfft // 2 + 1, nch, 2), dtype=torch.float32, device=x.device)
  if out.numel() == 0:
    return torch.view_as_complex(out)
  N.call('advoc_stft_f32', _ptr(x), b, nsamps, nch, nfft, nhop, 1 if pad_end else (0 if _lws_rule else 2), _ptr(win),
         _ptr(tw), _ptr(out), None, _stream())
  return torch.view_as_complex(out)


stft_batched = stft_tf


def magspec_tf(x, nfft, nhop, pad_end=True):
  """|stft_tf(x)| without materialising the complex spectrum (loader 'magspec' extract,
  reference: advoc/loader.py:117-121)."""
  if x.dim() != 4 or x.shape[2] != 1 or x.dtype != torch.float32:
    raise ValueError()
  x = x.contiguous()
  b, nsamps, _, nch = x.shape
  win, tw = _Consts.stft(nfft, nhop, x.device)
  frames = num_frames(nsamps, nfft, nhop, pad_end, tf_rule=True)
  out = torch.empty((b, frames, nfft // 2 + 1, nch), dtype=torch.float32, device=x.device)
  if out.numel() == 0:
    return out
  N.call('advoc_stft_f32', _ptr(x), b, nsamps, nch, nfft, nhop, 1 if pad_end else 2, _ptr(win),
         _ptr(tw), None, _ptr(out), _stream())
  return out


def waveform_to_melspec_tf(x, fs, nfft, nhop, mel_min=125, mel_max=7600, mel_num_bins=80,
                           norm_allow_clipping=True, norm_min_level_db=-100,
                           norm_ref_level_db=20):
  """x f32 [b, nsamps, 1, nch] (cuda) -> dB-normalised mel f32 [b, frames, mel_num_bins, nch].

  One fused kernel: frame -> window -> FFT -> |.| -> mel -> 20log10 -> clip.
  reference: advoc/spectral.py:158-227.
  """
  if x.dim() != 4 or x.shape[2] != 1:
    raise ValueError()
  if x.dtype != torch.float32:
    raise ValueError()
  if not norm_allow_clipping:
    raise NotImplementedError()  # advoc/spectral.py:220-223
  x = x.contiguous()
  b, nsamps, _, nch = x.shape
  win, tw = _Consts.stft(nfft, nhop, x.device)
  fb, ranges = _Consts.mel(fs, nfft, mel_min, mel_max, mel_num_bins, x.device)
  frames = num_frames(nsamps, nfft, nhop, True)
  out = torch.empty((b, frames, mel_num_bins, nch), dtype=torch.float32, device=x.device)
  if out.numel() == 0:
    return out
  N.call('advoc_melspec_f32', _ptr(x), b, nsamps, nch, nfft, nhop, _ptr(win), _ptr(tw), _ptr(fb),
         _ptr(ranges), mel_num_bins, float(norm_min_level_db), float(norm_ref_level_db),
         _ptr(out), _stream())
  return out


def waveform_to_r9y9_melspec_tf(x, fs=22050):
  """reference: advoc/spectral.py:272-291."""
  return waveform_to_melspec_tf(x, fs=fs, nfft=1024, nhop=256)


def matmul_lastdim(x, w, pow10_scale=False, min_level_db=-100., ref_level_db=20.):
  """y[..., n] = sum_k f(x[..., k]) * w[n, k] on the GPU (f = identity or the dB
  de-normalisation 10^((x*(-min_db)+min_db+ref_db)/20)).  reference: models/advoc/spectral_util.py:29-43,
  scripts/spectrogram_advoc.py:15-22, advoc/spectral.py:367-369."""
  if x.dtype != torch.float32 or w.dtype != torch.float32:
    raise ValueError()
  if w.dim() != 2 or x.shape[-1] != w.shape[1]:
    raise ValueError()
  x = x.contiguous()
  w = w.contiguous()
  rows = x.numel() // x.shape[-1] if x.numel() else 0
  y = torch.empty(x.shape[:-1] + (w.shape[0],), dtype=torch.float32, device=x.device)
  N.call('advoc_matmul_lastdim_f32', _ptr(x), _ptr(w), _ptr(y), rows, x.shape[-1], w.shape[0],
         1 if pow10_scale else 0, float(min_level_db), float(ref_level_db), _stream())
  return y


# ---------------------------------------------------------------------------
# numpy entry points (host buffers in, host buffers out) -- the reference's numpy API
# ---------------------------------------------------------------------------
def stft(x, nfft, nhop, pad_end=True):
  """x f32 [n, 1, 1] -> complex128 [frames, nfft//2+1, 1].  reference: advoc/spectral.py:11-41."""
  nsamps, nfeats, nch = x.shape
  if nfeats != 1:
    raise ValueError()
  if nch != 1:
    raise NotImplementedError('Can only take STFT of monaural signals')
  dev = _device()
  xd = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev).reshape(1, nsamps, 1, 1)
  X = stft_tf(xd, nfft, nhop, pad_end=pad_end, _lws_rule=True)
  return X[0].cpu().numpy().astype(np.complex128)


def waveform_to_melspec(x, fs, nfft, nhop, mel_min=125, mel_max=7600, mel_num_bins=80,
                        norm_allow_clipping=True, norm_min_level_db=-100,
                        norm_ref_level_db=20):
  """x f32 [n, 1, 1] -> f64 [frames, mel_num_bins, 1].  reference: advoc/spectral.py:98-154."""
  if x.dtype != np.float32:
    raise ValueError()
  nsamps, nfeats, nch = x.shape
  if nfeats != 1:
    raise ValueError()
  if nch != 1:
    raise NotImplementedError('Can only extract features from monaural signals')
  dev = _device()
  xd = torch.from_numpy(np.ascontiguousarray(x)).to(dev).reshape(1, nsamps, 1, 1)
  if not norm_allow_clipping:
    # the reference asserts on the un-clipped dB values (advoc/spectral.py:149-151)
    mag = magspec_tf(xd, nfft, nhop)[0, :, :, 0]
    fb, _ = _Consts.mel(fs, nfft, mel_min, mel_max, mel_num_bins, dev)
    mel = matmul_lastdim(mag, fb).cpu().numpy().astype(np.float64)
    min_level = np.exp(norm_min_level_db / 20 * np.log(10))
    db = 20 * np.log10(np.maximum(min_level, mel)) - norm_ref_level_db
    assert db.max() <= 0 and db.min() - norm_min_level_db >= 0
    return np.clip((db - norm_min_level_db) / -norm_min_level_db, 0, 1)[:, :, np.newaxis]
  out = waveform_to_melspec_tf(xd, fs, nfft, nhop, mel_min, mel_max, mel_num_bins, True,
                               norm_min_level_db, norm_ref_level_db)
  return out[0].cpu().numpy().astype(np.float64)


def waveform_to_tacotron2_melspec(x):
  """reference: advoc/spectral.py:230-247 (x at 24 kHz)."""
  return waveform_to_melspec(x, fs=24000, nfft=1200, nhop=300, norm_min_level_db=-40)


def waveform_to_r9y9_melspec(x, fs=22050):
  """reference: advoc/spectral.py:250-269."""
  return waveform_to_melspec(x, fs=fs, nfft=1024, nhop=256)


# ---------------------------------------------------------------------------
# inversion (SURVEY.md section 8(f) "next" row): ISTFT and Griffin-Lim on the GPU
# ---------------------------------------------------------------------------
def istft_tf(X, nfft, nhop):
  """Batched inverse STFT.  X c64 [b, frames, nfft//2+1] (cuda) -> f32 [b, (frames-1)*nhop + nfft].
  lws istft with perfectrec=False (synthesis window == analysis window).
  reference: advoc/spectral.py:303,307,322 (`lws_proc.istft`)."""
  if X.dim() != 3 or X.shape[2] != nfft // 2 + 1 or X.dtype != torch.complex64:
    raise ValueError()
  b, frames, _ = X.shape
  win, tw = _Consts.stft(nfft, nhop, X.device)
  nout = (frames - 1) * nhop + nfft if frames > 0 else 0
  out = torch.zeros((b, nout), dtype=torch.float32, device=X.device)
  if out.numel() == 0:
    return out
  spec = torch.view_as_real(X.contiguous())
  fr = torch.empty((b, frames, nfft), dtype=torch.float32, device=X.device)
  N.call('advoc_istft_frames_f32', _ptr(spec), b, frames, nfft, nhop, _ptr(win), _ptr(tw), _ptr(fr),
         _stream())
  N.call('advoc_overlap_add_f32', _ptr(fr), b, frames, nfft, nhop, _ptr(out), _stream())
  return out


def griffin_lim_tf(X_mag, nfft, nhop, ngl=60, init_phase=None, generator=None):
  """Batched Griffin-Lim.  X_mag f32 [b, frames, nfft//2+1] (cuda) -> f32 [b, (frames-1)*nhop+nfft].
  Each iteration is two kernels (fused stft -> phase -> istft frames, then overlap-add).
  `init_phase` (radians, same shape) replaces the reference's `2*pi*np.random.rand` start."""
  if X_mag.dim() != 3 or X_mag.shape[2] != nfft // 2 + 1 or X_mag.dtype != torch.float32:
    raise ValueError()
  b, frames, bins = X_mag.shape
  mag = X_mag.abs().contiguous()
  if init_phase is None:
    init_phase = 2 * np.pi * torch.rand(mag.shape, device=mag.device, generator=generator)
  x = istft_tf(torch.polar(mag, init_phase.to(torch.float32)), nfft, nhop)
  if frames == 0:
    return x
  win, tw = _Consts.stft(nfft, nhop, mag.device)
  fr = torch.empty((b, frames, nfft), dtype=torch.float32, device=mag.device)
  for _ in range(ngl):
    N.call('advoc_griffin_lim_iter_f32', _ptr(x), x.shape[1], _ptr(mag), b, frames, nfft, nhop,
           _ptr(win), _ptr(tw), _ptr(fr), _stream())
    N.call('advoc_overlap_add_f32', _ptr(fr), b, frames, nfft, nhop, _ptr(x), _stream())
  return x


def magspec_to_waveform_griffin_lim(X_mag, nfft, nhop, ngl=60, init_phase=None):
  """X_mag [frames, bins, 1] -> f32 [n, 1, 1].  reference: advoc/spectral.py:294-311."""
  nsamps, nbins, nch = X_mag.shape
  if nch != 1:
    raise NotImplementedError('Can only invert monaural signals')
  dev = _device()
  mag = torch.from_numpy(np.ascontiguousarray(np.abs(X_mag[:, :, 0]), dtype=np.float32)).to(dev)[None]
  ph = None
  if init_phase is not None:
    ph = torch.from_numpy(np.ascontiguousarray(init_phase, dtype=np.float32)).to(dev)[None]
  x = griffin_lim_tf(mag, nfft, nhop, ngl=ngl, init_phase=ph)
  return x[0].cpu().numpy()[:, np.newaxis, np.newaxis].astype(np.float32)


def lws_tf(X_mag, nfft, nhop, iterations=100, init_phase=None, generator=None):
  """Batched LWS-style phase reconstruction.  X_mag f32 [b, frames, nfft//2+1] (cuda) -> complex
  spectrogram c64 of the same shape with the given magnitudes.

  Le Roux et al.'s local weighted sums ("Fast signal reconstruction from magnitude STFT spectrogram
  based on spectrogram consistency", DAFx 2010; the `lws` package's batch stage) update every bin's
  phase from the consistency operator F = STFT o ISTFT applied to its NEIGHBOURS, leaving out the bin's
  own contribution alpha_0(0) * X (which only slows convergence down):
      X <- |X_mag| * phase( F(X) - alpha_0(0) * X ),   alpha_0(0) = sum_k w(k)^2 / nfft = nhop / nfft
  for this window.  `lws` truncates the weights to a (2L+1) x (2Q-1) neighbourhood and skips
  low-magnitude bins to save CPU time; on the GPU the untruncated operator is two kernels (the istft /
  stft pair of this module), so it is applied in full.  The reference's `run_lws(mode='speech')` also
  runs its own frame-sequential initialisation whose source is not part of the reference: here the
  start is `init_phase` or uniform random phase.  **Parity unpinned** (no lws build, no reproducible
  golden: SURVEY.md section 8(f) row 1)."""
  if X_mag.dim() != 3 or X_mag.shape[2] != nfft // 2 + 1 or X_mag.dtype != torch.float32:
    raise ValueError()
  mag = X_mag.abs().contiguous()
  if init_phase is None:
    init_phase = 2 * np.pi * torch.rand(mag.shape, device=mag.device, generator=generator)
  X = torch.polar(mag, init_phase.to(torch.float32))
  if mag.shape[1] == 0:
    return X
  c0 = float(nhop) / float(nfft)
  for _ in range(iterations):
    x = istft_tf(X, nfft, nhop)
    Y = stft_tf(x[:, :, None, None], nfft, nhop, pad_end=False)[:, :, :, 0]
    Z = Y - c0 * X
    X = torch.polar(mag, torch.angle(Z))
  return X


def magspec_to_waveform_lws(X_mag, nfft, nhop, iterations=100, init_phase=None):
  """X_mag [frames, bins, 1] -> f32 [n, 1, 1].  reference: advoc/spectral.py:314-326
  (`lws.lws(nfft, nhop, mode='speech', perfectrec=False).run_lws` then `.istft`); phase estimate: `lws_tf`
  (the batch LWS iteration with untruncated weights; parity unpinned)."""
  nsamps, nbins, nch = X_mag.shape
  if nch != 1:
    raise NotImplementedError('Can only invert monaural signals')
  dev = _device()
  mag = torch.from_numpy(np.ascontiguousarray(np.abs(X_mag[:, :, 0]), dtype=np.float32)).to(dev)[None]
  ph = None
  if init_phase is not None:
    ph = torch.from_numpy(np.ascontiguousarray(init_phase, dtype=np.float32)).to(dev)[None]
  x = istft_tf(lws_tf(mag, nfft, nhop, iterations=iterations, init_phase=ph), nfft, nhop)
  return x[0].cpu().numpy()[:, np.newaxis, np.newaxis].astype(np.float32)


def melspec_to_waveform(X_mel_dbnorm, fs, nfft, nhop, mel_min=125, mel_max=7600,
                        norm_min_level_db=-100, norm_ref_level_db=20, phase_estimation='lws',
                        waveform_len=None):
  """dB-normalised mel f64 [frames, mels, 1] -> waveform f32 [n, 1, 1]: dB de-normalise, pinv mel
  (one fused kernel), clamp at 0, phase estimation.  reference: advoc/spectral.py:330-395."""
  if X_mel_dbnorm.dtype != np.float64:
    raise ValueError()
  nsamps, mel_num_bins, nch = X_mel_dbnorm.shape
  if nch != 1:
    raise NotImplementedError('Can only invert monaural signals')
  if phase_estimation != 'lws':
    if phase_estimation[:2] != 'gl':
      raise ValueError()
    try:
      ngl = int(phase_estimation[2:])
    except Exception:
      raise ValueError()
  dev = _device()
  inv = _Consts.inv_mel(fs, nfft, mel_min, mel_max, mel_num_bins, dev)
  mel = torch.from_numpy(np.ascontiguousarray(X_mel_dbnorm[:, :, 0], dtype=np.float32)).to(dev)
  X_mag = torch.clamp_min(matmul_lastdim(mel, inv, pow10_scale=True, min_level_db=norm_min_level_db,
                                         ref_level_db=norm_ref_level_db), 0.)
  if phase_estimation == 'lws':
    x = magspec_to_waveform_lws(X_mag.cpu().numpy()[:, :, np.newaxis], nfft, nhop)
  else:
    x = griffin_lim_tf(X_mag[None], nfft, nhop, ngl=ngl)[0].cpu().numpy()[:, np.newaxis, np.newaxis]
  if waveform_len is not None:
    x_len = x.shape[0]
    if x_len < waveform_len:
      x = np.pad(x, [[0, waveform_len - x_len], [0, 0], [0, 0]], 'constant')
    elif x_len > waveform_len:
      x = x[:waveform_len]
  return x.astype(np.float32)


def r9y9_melspec_to_waveform(X_mel_dbnorm, fs=22050, phase_estimation='lws', waveform_len=None):
  """reference: advoc/spectral.py:398-420."""
  return melspec_to_waveform(X_mel_dbnorm, fs=fs, nfft=1024, nhop=256,
                             phase_estimation=phase_estimation, waveform_len=waveform_len)
