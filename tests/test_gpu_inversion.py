"""GPU parity of the inversion kernels (ISTFT, Griffin-Lim) against the numpy oracle
(oracle/spectral_np.py: lws istft / the reference's GL loop restated).  The reference's own
inversion goldens (tests/test_spectral.py:156-208) need librosa resampling and lws.run_lws and
are not reproducible here, so this path is "parity unpinned"; tolerance 1e-4 relative L2."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(a, b):
  a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
  return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def test_istft_matches_oracle_and_inverts_stft():
  import torch
  from advoc_b200 import spectral as S
  from oracle import spectral_np as O
  rng = np.random.RandomState(0)
  X = (rng.randn(2, 37, 513) + 1j * rng.randn(2, 37, 513)).astype(np.complex64)
  got = S.istft_tf(torch.from_numpy(X).cuda(), 1024, 256).cpu().numpy()
  for b in range(2):
    ref = O.istft(X[b].astype(np.complex128), 1024, 256)
    assert got[b].shape == ref.shape == (36 * 256 + 1024,)
    assert _rel(got[b], ref) < 1e-4
  # stft -> istft reproduces the interior of a signal (window is power-complementary at 75 % overlap)
  x = rng.uniform(-1, 1, (1, 16384, 1, 1)).astype(np.float32)
  Xs = S.stft_tf(torch.from_numpy(x).cuda(), 1024, 256, pad_end=False)[:, :, :, 0]
  y = S.istft_tf(Xs.contiguous(), 1024, 256).cpu().numpy()[0]
  assert y.shape == (16384,)
  assert _rel(y[1024:-1024], x[0, 1024:-1024, 0, 0]) < 1e-4


def test_griffin_lim_matches_oracle_with_injected_phase(golden_dir):
  from advoc_b200 import spectral as S
  from oracle import spectral_np as O
  mel = np.load(os.path.join(golden_dir, 'mono_22k_r9y9_mel.npy')).T[:64]          # [64, 80]
  Winv = O.create_inverse_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80)
  X_mag = np.maximum(0., O.tacotron_mel_to_mag(mel, Winv))[:, :, np.newaxis]       # [64, 513, 1]
  phase = 2 * np.pi * np.random.RandomState(1).rand(64, 513)

  class _Rng(object):          # makes the oracle start from the same phase
    def rand(self, *shape):
      return phase / (2 * np.pi)

  for ngl in (0, 5):
    ref = O.griffin_lim(X_mag, 1024, 256, ngl=ngl, rng=_Rng())
    got = S.magspec_to_waveform_griffin_lim(X_mag, 1024, 256, ngl=ngl, init_phase=phase)
    assert got.shape == ref.shape == (63 * 256 + 1024, 1, 1) and got.dtype == np.float32
    assert _rel(got, ref) < 1e-4 * (1 + ngl)      # GL is a fixed-point iteration: errors compound mildly


def test_melspec_to_waveform_contract(golden_dir):
  from advoc_b200 import spectral as S
  mel = np.load(os.path.join(golden_dir, 'mono_22k_r9y9_mel.npy')).T[:, :, np.newaxis].copy()  # [325,80,1]
  x = S.r9y9_melspec_to_waveform(mel, phase_estimation='gl10', waveform_len=82432)
  assert x.shape == (82432, 1, 1) and x.dtype == np.float32 and np.isfinite(x).all()
  assert 0.005 < float(np.abs(x).mean()) < 0.5
  with pytest.raises(ValueError):
    S.melspec_to_waveform(mel.astype(np.float32), 22050, 1024, 256)
  with pytest.raises(NotImplementedError):
    S.melspec_to_waveform(np.zeros((10, 80, 2)), 22050, 1024, 256)
  with pytest.raises(ValueError):
    S.melspec_to_waveform(mel, 22050, 1024, 256, phase_estimation='foo')
  x = S.r9y9_melspec_to_waveform(mel[:64])   # default phase_estimation='lws' (advoc/spectral.py:398)
  assert x.shape == (63 * 256 + 1024, 1, 1) and x.dtype == np.float32 and np.isfinite(x).all()
  # tacotron2-style normalisation (-40 dB floor, advoc/spectral.py:230-247) inverts too
  y = S.melspec_to_waveform(mel[:32], 22050, 1024, 256, norm_min_level_db=-40, phase_estimation='gl2')
  assert y.shape == (31 * 256 + 1024, 1, 1) and np.isfinite(y).all()


def test_lws_matches_oracle_and_beats_griffin_lim(golden_dir):
  """`magspec_to_waveform_lws` (batch LWS iteration, untruncated weights; parity unpinned): equals the
  numpy restatement from the same initial phase, and after the same number of iterations its
  spectrogram is more consistent with the target magnitudes than Griffin-Lim's (the point of leaving
  the centre term out)."""
  from advoc_b200 import spectral as S
  from oracle import spectral_np as O
  mel = np.load(os.path.join(golden_dir, 'mono_22k_r9y9_mel.npy')).T[:64]
  Winv = O.create_inverse_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80)
  X_mag = np.maximum(0., O.tacotron_mel_to_mag(mel, Winv))[:, :, np.newaxis]
  phase = 2 * np.pi * np.random.RandomState(1).rand(64, 513)

  class _Rng(object):
    def rand(self, *shape):
      return phase / (2 * np.pi)

  ref = O.lws(X_mag, 1024, 256, iterations=5, rng=_Rng())
  got = S.magspec_to_waveform_lws(X_mag, 1024, 256, iterations=5, init_phase=phase)
  assert got.shape == ref.shape == (63 * 256 + 1024, 1, 1) and got.dtype == np.float32
  assert _rel(got, ref) < 2e-3

  def inconsistency(x):
    mag = np.abs(O.stft(x[:, 0, 0][:, None, None], 1024, 256, pad_end=False)[:, :, 0])
    return float(np.linalg.norm(mag[:64] - X_mag[:, :, 0]) / np.linalg.norm(X_mag))

  x_lws = S.magspec_to_waveform_lws(X_mag, 1024, 256, iterations=30, init_phase=phase)
  x_gl = S.magspec_to_waveform_griffin_lim(X_mag, 1024, 256, ngl=30, init_phase=phase)
  assert inconsistency(x_lws) < inconsistency(x_gl)
