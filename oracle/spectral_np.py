"""CPU restatement (numpy) of the reference's spectral feature path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product path (`advoc_b200/`) never does.

The reference implements this path on top of three third-party packages that
are absent from this image and not vendored in `/root/reference`:
  * lws==1.2            (setup.py:19)  -- STFT / window     (C++/Cython)
  * librosa==0.6.3      (setup.py:18)  -- Slaney mel filterbank
  * tensorflow<=1.13.1  (setup.py:17)  -- `tf.contrib.signal.stft`, matmul, log
so every function below restates the *published* algorithm of the dependency
and is anchored on the reference's own call sites and golden numbers
(`tests/test_spectral.py`), which `tests/test_oracle_spectral.py` re-checks.

Parity status: PINNED for window + framing + rfft (test_spectral.py:38-46) and
for the Slaney filterbank + f32 dB normalisation (test_spectral.py:123-139).
Not reproducible here: any golden that goes through librosa's kaiser_best
resampler.
"""
import numpy as np


# ---------------------------------------------------------------------------
# window  (reference: advoc/spectral.py:44-57; lws.hann(N, symmetric=True,
# use_offset=False) then sqrt(hann * 2 * hop / N))
# ---------------------------------------------------------------------------
def lws_hann(nfft):
  """lws 1.2 `hann(N, symmetric=True, use_offset=False)`: half-sample-offset Hann."""
  i = np.arange(nfft, dtype=np.float64)
  return 0.5 * (1.0 - np.cos(2.0 * np.pi * (i + 0.5) / nfft))


def lws_hann_default(nfft, nhop, dtype=np.float64):
  """Analysis window of `lws.lws(nfft, nhop)`; advoc/spectral.py:55-56."""
  return np.sqrt(lws_hann(nfft) * 2.0 * nhop / nfft).astype(dtype)


# ---------------------------------------------------------------------------
# framing + STFT  (reference: advoc/spectral.py:11-41 and :60-83)
# ---------------------------------------------------------------------------
def num_frames(nsamps, nfft, nhop, pad_end=True):
  """Frame count rule.

  pad_end=True : ceil(n/hop) (advoc/spectral.py:33); the tail is zero padded so
                 the last frame is complete (:35-39).
  pad_end=False: lws itself still zero-pads the last partial frame:
                 ceil((n - nfft)/hop) + 1 (tests/test_spectral.py:35-36: 16000 -> 60).
  """
  if nsamps <= 0:
    return 0
  if pad_end:
    return int(np.ceil(float(nsamps) / nhop) + 1e-6)
  return max(int(np.ceil(float(nsamps - nfft) / nhop)) + 1, 1)


def frame_signal(x, nfft, nhop, pad_end=True):
  """[n] -> [frames, nfft], zero padded tail; no centering, no reflection."""
  n = x.shape[0]
  m = num_frames(n, nfft, nhop, pad_end)
  need = (m - 1) * nhop + nfft if m > 0 else 0
  if need > n:
    x = np.concatenate([x, np.zeros(need - n, dtype=x.dtype)])
  idx = np.arange(m)[:, None] * nhop + np.arange(nfft)[None, :]
  return x[idx] if m > 0 else np.zeros((0, nfft), dtype=x.dtype)


def stft(x, nfft, nhop, pad_end=True):
  """advoc/spectral.py:11-41.  x f32 [n,1,1] -> c128 [frames, nfft//2+1, 1]."""
  nsamps, nfeats, nch = x.shape
  if nfeats != 1:
    raise ValueError()
  if nch != 1:
    raise NotImplementedError('Can only take STFT of monaural signals')
  xs = x[:, 0, 0].astype(np.float64)
  frames = frame_signal(xs, nfft, nhop, pad_end)
  win = lws_hann_default(nfft, nhop, np.float64)
  X = np.fft.rfft(frames * win[None, :], axis=1)
  return X.astype(np.complex128)[:, :, np.newaxis]


def stft_f32(x, nfft, nhop, pad_end=True):
  """advoc/spectral.py:60-83 (`stft_tf`).  x f32 [b,n,1,ch] -> c64 [b,frames,bins,ch].

  tf.contrib.signal.stft: frame (pad_end zero pads), multiply by the f32 window,
  rfft with fft_length=nfft, all in float32.
  """
  b, nsamps, nfeats, nch = x.shape
  if nfeats != 1:
    raise ValueError()
  if x.dtype != np.float32:
    raise ValueError()
  win = lws_hann_default(nfft, nhop, np.float64).astype(np.float32)
  m = num_frames(nsamps, nfft, nhop, pad_end)
  out = np.zeros((b, m, nfft // 2 + 1, nch), dtype=np.complex64)
  for bi in range(b):
    for c in range(nch):
      fr = frame_signal(x[bi, :, 0, c], nfft, nhop, pad_end) * win[None, :]
      # scipy/numpy pocketfft keeps float32 when asked through scipy.fft
      try:
        import scipy.fft as sfft
        X = sfft.rfft(fr.astype(np.float32), axis=1)
      except ImportError:  # pragma: no cover
        X = np.fft.rfft(fr, axis=1)
      out[bi, :, :, c] = X.astype(np.complex64)
  return out


# ---------------------------------------------------------------------------
# Slaney mel filterbank (librosa 0.6.3 `filters.mel(sr, n_fft, n_mels, fmin,
# fmax, htk=False, norm=1)`; reference call site advoc/spectral.py:86-88)
# ---------------------------------------------------------------------------
_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
  f = np.asarray(f, dtype=np.float64)
  lin = f / _F_SP
  log = _MIN_LOG_MEL + np.log(np.maximum(f, 1e-300) / _MIN_LOG_HZ) / _LOGSTEP
  return np.where(f >= _MIN_LOG_HZ, log, lin)


def mel_to_hz(m):
  m = np.asarray(m, dtype=np.float64)
  lin = m * _F_SP
  log = _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL))
  return np.where(m >= _MIN_LOG_MEL, log, lin)


def create_mel_filterbank(fs, nfft, fmin=0.0, fmax=None, n_mels=128):
  """f64 [n_mels, nfft//2+1]; area-normalised triangles on the Slaney scale."""
  if fmax is None:
    fmax = fs / 2.0
  nbins = nfft // 2 + 1
  fftfreqs = np.linspace(0.0, fs / 2.0, nbins)
  edges = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
  W = np.zeros((n_mels, nbins), dtype=np.float64)
  for i in range(n_mels):
    lo, ce, hi = edges[i], edges[i + 1], edges[i + 2]
    rise = (fftfreqs - lo) / (ce - lo)
    fall = (hi - fftfreqs) / (hi - ce)
    W[i] = np.maximum(0.0, np.minimum(rise, fall)) * (2.0 / (hi - lo))
  return W


def create_inverse_mel_filterbank(fs, nfft, fmin=0.0, fmax=None, n_mels=128):
  """advoc/spectral.py:91-94: Moore-Penrose pseudo-inverse, f64 [nbins, n_mels]."""
  return np.linalg.pinv(create_mel_filterbank(fs, nfft, fmin=fmin, fmax=fmax, n_mels=n_mels))


# ---------------------------------------------------------------------------
# waveform -> dB-normalised mel  (advoc/spectral.py:98-154 f64, :158-227 f32)
# ---------------------------------------------------------------------------
def waveform_to_melspec(x, fs, nfft, nhop, mel_min=125, mel_max=7600, mel_num_bins=80,
                        norm_allow_clipping=True, norm_min_level_db=-100, norm_ref_level_db=20):
  if x.dtype != np.float32:
    raise ValueError()
  nsamps, nfeats, nch = x.shape
  if nfeats != 1:
    raise ValueError()
  if nch != 1:
    raise NotImplementedError('Can only extract features from monaural signals')
  X_mag = np.abs(stft(x, nfft, nhop)[:, :, 0])
  W = create_mel_filterbank(fs, nfft, fmin=mel_min, fmax=mel_max, n_mels=mel_num_bins)
  X_mel = np.dot(W, X_mag.T).T
  min_level = np.exp(norm_min_level_db / 20 * np.log(10))
  X_db = 20 * np.log10(np.maximum(min_level, X_mel)) - norm_ref_level_db
  if not norm_allow_clipping:
    assert X_db.max() <= 0 and X_db.min() - norm_min_level_db >= 0
  return np.clip((X_db - norm_min_level_db) / -norm_min_level_db, 0, 1)[:, :, np.newaxis]


def waveform_to_melspec_f32(x, fs, nfft, nhop, mel_min=125, mel_max=7600, mel_num_bins=80,
                            norm_allow_clipping=True, norm_min_level_db=-100,
                            norm_ref_level_db=20):
  """f32 batched twin (`waveform_to_melspec_tf`).  x [b,n,1,ch] -> [b,frames,mels,ch]."""
  b, nsamps, one, nch = x.shape
  if one != 1:
    raise ValueError()
  if x.dtype != np.float32:
    raise ValueError()
  if not norm_allow_clipping:
    raise NotImplementedError()
  X_mag = np.abs(stft_f32(x, nfft, nhop)).astype(np.float32)        # [b,t,f,ch]
  W = create_mel_filterbank(fs, nfft, fmin=mel_min, fmax=mel_max,
                            n_mels=mel_num_bins).astype(np.float32)
  X_mel = np.einsum('btfc,mf->btmc', X_mag, W).astype(np.float32)
  min_level = np.float32(np.exp(norm_min_level_db / 20 * np.log(10)))
  log10 = np.log(np.maximum(min_level, X_mel)) / np.log(np.float32(10))
  X_db = (np.float32(20) * log10 - np.float32(norm_ref_level_db)).astype(np.float32)
  out = (X_db - np.float32(norm_min_level_db)) / np.float32(-norm_min_level_db)
  return np.clip(out, 0, 1).astype(np.float32)


def waveform_to_r9y9_melspec(x, fs=22050):
  return waveform_to_melspec(x, fs=fs, nfft=1024, nhop=256)


def waveform_to_r9y9_melspec_f32(x, fs=22050):
  return waveform_to_melspec_f32(x, fs=fs, nfft=1024, nhop=256)


def waveform_to_tacotron2_melspec(x):
  return waveform_to_melspec(x, fs=24000, nfft=1200, nhop=300, norm_min_level_db=-40)


# ---------------------------------------------------------------------------
# magnitude STFT batched (loader 'magspec' extract: |stft_tf|), linear mel and its
# pseudo-inverse lift (models/advoc/spectral_util.py:29-43), dB-denorm + pinv
# (scripts/spectrogram_advoc.py:15-22)
# ---------------------------------------------------------------------------
def magspec_f32(x, nfft, nhop, pad_end=True):
  return np.abs(stft_f32(x, nfft, nhop, pad_end)).astype(np.float32)


def mag_to_mel_linear_spec(mag, W):
  """spectral_util.py:29-32.  mag [b,t,513,1] f32, W [80,513] -> [b,t,80,1]."""
  return np.tensordot(mag[:, :, :, 0], W.astype(np.float32).T, axes=1)[..., np.newaxis]


def mel_linear_to_mag_spec(mel, Winv):
  """spectral_util.py:34-43.  mel [b,t,80,1], Winv [513,80] -> [b,t,513,1]; no clamp."""
  return np.tensordot(mel[:, :, :, 0], Winv.astype(np.float32).T, axes=1)[..., np.newaxis]


def tacotron_mel_to_mag(X_mel_dbnorm, Winv):
  """scripts/spectrogram_advoc.py:15-22 (f64).  [T,80] -> [T,513]."""
  X_db = X_mel_dbnorm * 100.0 - 100.0
  X_mel = np.power(10, (X_db + 20.0) / 20)
  return np.dot(X_mel, Winv.T)


# ---------------------------------------------------------------------------
# inverse STFT / Griffin-Lim ("next" row; lws istft with perfectrec=False:
# overlap-add of irfft(frames) * synthesis window, synthesis window == analysis
# window, whose squared 75%-overlap sum is 1).  advoc/spectral.py:294-311.
# ---------------------------------------------------------------------------
def istft(X, nfft, nhop):
  """X c128 [frames, bins] -> f64 [(frames-1)*hop + nfft]."""
  m = X.shape[0]
  win = lws_hann_default(nfft, nhop, np.float64)
  fr = np.fft.irfft(X, n=nfft, axis=1) * win[None, :]
  out = np.zeros((m - 1) * nhop + nfft if m > 0 else 0, dtype=np.float64)
  for i in range(m):
    out[i * nhop:i * nhop + nfft] += fr[i]
  return out


def griffin_lim(X_mag, nfft, nhop, ngl=60, rng=None):
  """advoc/spectral.py:294-311 with lws stft/istft restated.  X_mag [frames,bins,1]."""
  if X_mag.shape[2] != 1:
    raise NotImplementedError('Can only invert monaural signals')
  rng = np.random if rng is None else rng
  mag = np.abs(X_mag[:, :, 0]).astype(np.complex128)
  angles = np.exp(2j * np.pi * rng.rand(*mag.shape))
  x = istft(mag * angles, nfft, nhop)
  win = lws_hann_default(nfft, nhop, np.float64)
  for _ in range(ngl):
    fr = frame_signal(x, nfft, nhop, pad_end=False)
    S = np.fft.rfft(fr * win[None, :], axis=1)
    angles = np.exp(1j * np.angle(S))
    x = istft(mag * angles, nfft, nhop)
  return x[:, np.newaxis, np.newaxis].astype(np.float32)


def lws(X_mag, nfft, nhop, iterations=100, rng=None):
  """Batch local-weighted-sums phase reconstruction with untruncated weights (Le Roux et al., DAFx
  2010): X <- |X| * phase(STFT(ISTFT(X)) - alpha_0(0) * X), alpha_0(0) = nhop / nfft for the lws window.
  Restates the product's `lws_tf` (advoc/spectral.py:314-326 calls `lws.run_lws`, whose source and
  goldens are not available: parity unpinned).  X_mag [frames, bins, 1] -> waveform f32 [n, 1, 1]."""
  if X_mag.shape[2] != 1:
    raise NotImplementedError('Can only invert monaural signals')
  rng = np.random if rng is None else rng
  mag = np.abs(X_mag[:, :, 0]).astype(np.float64)
  X = mag * np.exp(2j * np.pi * rng.rand(*mag.shape))
  win = lws_hann_default(nfft, nhop, np.float64)
  c0 = float(nhop) / float(nfft)
  for _ in range(iterations):
    x = istft(X, nfft, nhop)
    Y = np.fft.rfft(frame_signal(x, nfft, nhop, pad_end=False) * win[None, :], axis=1)
    Z = Y - c0 * X
    X = mag * np.exp(1j * np.angle(Z))
  return istft(X, nfft, nhop)[:, np.newaxis, np.newaxis].astype(np.float32)
