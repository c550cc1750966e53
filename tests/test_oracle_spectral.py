"""Pins the numpy oracle (oracle/spectral_np.py) to the reference's own golden numbers.

Every constant below is copied from /root/reference/tests/test_spectral.py (file:line cited);
the wav fixture tests/golden/sc09.wav is the reference's tests/audio/sc09.wav and
tests/golden/mono_22k_r9y9_mel.npy is its tests/audio/mono_22k_r9y9.pkl re-saved as .npy
(tests/golden/make_golden.py).  CPU only.
"""
import os

import numpy as np
import pytest
from scipy.io import wavfile

from oracle import spectral_np as O


def _load_sc09(golden_dir):
  fs, x = wavfile.read(os.path.join(golden_dir, 'sc09.wav'))
  assert fs == 16000 and x.dtype == np.int16 and x.shape == (16000,)
  return (x.astype(np.float32) / 32768.0).reshape(-1, 1, 1)


def test_stft_shapes_and_sums(golden_dir):
  # tests/test_spectral.py:27-46
  x = _load_sc09(golden_dir)
  X = O.stft(x, 1024, 256)
  assert X.dtype == np.complex128 and X.shape == (63, 513, 1)        # :31-33
  assert O.stft(x, 1024, 256, pad_end=False).shape == (60, 513, 1)   # :35-36
  xp = np.pad(x, [[0, 384], [0, 0], [0, 0]], 'constant')             # :38
  X = O.stft(xp, 1024, 256)
  assert X.shape == (64, 513, 1)                                     # :40
  mag = np.abs(X)
  assert round(abs(float(mag.sum()) - 2148.69), 2) == 0              # :42
  assert round(abs(float(mag[33].sum()) - 55.45), 2) == 0            # :43
  assert round(abs(float(mag[40].sum()) - 20.35), 2) == 0            # :44


def test_stft_f32_matches_f64(golden_dir):
  # tests/test_spectral.py:49-76 (`stft_tf` == lws stft on sc09, batch item 1)
  x = _load_sc09(golden_dir)
  xp = np.pad(x, [[0, 384], [0, 0], [0, 0]], 'constant')
  X32 = O.stft_f32(xp[np.newaxis].astype(np.float32), 1024, 256)
  assert X32.dtype == np.complex64 and X32.shape == (1, 64, 513, 1)
  mag = np.abs(X32[0])
  assert round(abs(float(mag.sum()) - 2148.69), 2) == 0
  assert round(abs(float(mag[33].sum()) - 55.45), 2) == 0
  assert round(abs(float(mag[40].sum()) - 20.35), 2) == 0
  X64 = O.stft(xp, 1024, 256)
  assert np.abs(X32[0] - X64).max() < 1e-4


def test_r9y9_f32_noise_goldens():
  # tests/test_spectral.py:123-139: seeded uniform noise through the f32 mel path
  np.random.seed(0)
  n1 = np.random.uniform(-1, 1, (1, 82432, 1, 1)).astype(np.float32)
  n2 = np.random.uniform(-1, 1, (1, 82432, 1, 2)).astype(np.float32)
  m2 = O.waveform_to_r9y9_melspec_f32(n2)
  m1 = O.waveform_to_r9y9_melspec_f32(n1)
  assert m2.shape == (1, 322, 80, 2) and m2.dtype == np.float32
  # the reference asserts 3 decimal places on a float32 sum of 25760 values (ulp there is
  # 2e-3); summation order differs between TF and numpy, so allow two ulps
  assert abs(float(m2[0, :, :, 0].sum()) - 18328.508) < 4e-3
  assert abs(float(m2[0, :, :, 1].sum()) - 18332.746) < 4e-3
  assert abs(float(m1[0, :, :, 0].sum()) - 18319.934) < 4e-3


def test_r9y9_pickle_fixture(golden_dir):
  # tests/test_spectral.py:98-106,140: the r9y9 reference mel; sum of frames [3:] == 5121.489...
  mel = np.load(os.path.join(golden_dir, 'mono_22k_r9y9_mel.npy'))
  assert mel.shape == (80, 325) and mel.dtype == np.float64
  assert abs(float(mel.T[3:].sum()) - 5121.489431680474) < 1e-6


def test_frame_count_rule():
  # advoc/spectral.py:32-39 and tests/test_spectral.py:31-36
  assert O.num_frames(16000, 1024, 256, True) == 63
  assert O.num_frames(16000, 1024, 256, False) == 60
  assert O.num_frames(16384, 1024, 256, True) == 64
  assert O.num_frames(0, 1024, 256, True) == 0
  assert O.num_frames(82432, 1024, 256, True) == 322
  assert O.num_frames(22050, 1024, 256, True) == 87


def test_window_is_power_complementary():
  # the 75%-overlap squared sum of the lws window is 1 -> istft(stft(x)) == x in the interior
  w = O.lws_hann_default(1024, 256)
  s = (w.reshape(4, 256) ** 2).sum(axis=0)
  assert np.allclose(s, 1.0, atol=1e-12)


def test_mel_filterbank_properties():
  W = O.create_mel_filterbank(22050, 1024, fmin=125, fmax=7600, n_mels=80)
  assert W.shape == (80, 513) and W.dtype == np.float64
  assert (W >= 0).all() and (W.sum(axis=1) > 0).all()
  Winv = O.create_inverse_mel_filterbank(22050, 1024, fmin=125, fmax=7600, n_mels=80)
  assert Winv.shape == (513, 80)
  assert np.allclose(W @ Winv, np.eye(80), atol=1e-8)


def test_error_contracts():
  # advoc/spectral.py:24-27,131-138
  with pytest.raises(ValueError):
    O.stft(np.zeros((10, 2, 1), np.float32), 1024, 256)
  with pytest.raises(NotImplementedError):
    O.stft(np.zeros((10, 1, 2), np.float32), 1024, 256)
  with pytest.raises(ValueError):
    O.waveform_to_melspec(np.zeros((10, 1, 1), np.float64), 22050, 1024, 256)
