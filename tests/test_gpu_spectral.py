"""GPU parity of the spectral feature kernels (csrc/spectral.cu, through the C-ABI) against
the numpy oracle and the reference's golden numbers.  Tolerance: 1e-4 relative L2
(BASELINE.json north_star) -- written in each assert."""
import os

import numpy as np
import pytest
from scipy.io import wavfile

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _rel(a, b):
  a = np.asarray(a)
  b = np.asarray(b)
  return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def _sc09(golden_dir):
  _, x = wavfile.read(os.path.join(golden_dir, 'sc09.wav'))
  return (x.astype(np.float32) / 32768.0).reshape(-1, 1, 1)


def test_stft_reference_goldens(golden_dir):
  # reference tests/test_spectral.py:27-46, through the numpy-API drop-in
  from advoc_b200 import spectral as S
  x = _sc09(golden_dir)
  X = S.stft(x, 1024, 256)
  assert X.dtype == np.complex128 and X.shape == (63, 513, 1)
  assert S.stft(x, 1024, 256, pad_end=False).shape == (60, 513, 1)
  xp = np.pad(x, [[0, 384], [0, 0], [0, 0]], 'constant')
  X = S.stft(xp, 1024, 256)
  assert X.shape == (64, 513, 1)
  mag = np.abs(X)
  assert round(abs(float(mag.sum()) - 2148.69), 2) == 0
  assert round(abs(float(mag[33].sum()) - 55.45), 2) == 0
  assert round(abs(float(mag[40].sum()) - 20.35), 2) == 0


def test_stft_matches_oracle(golden_dir):
  from advoc_b200 import spectral as S
  from oracle import spectral_np as O
  x = _sc09(golden_dir)
  for pad_end in (True, False):
    got = S.stft(x, 1024, 256, pad_end=pad_end)
    ref = O.stft(x, 1024, 256, pad_end=pad_end)
    assert got.shape == ref.shape
    assert _rel(got, ref) < TOL


def test_stft_tf_batched_multichannel():
  import torch
  from advoc_b200 import spectral as S
  from oracle import spectral_np as O
  rng = np.random.RandomState(3)
  x = rng.uniform(-1, 1, (3, 5000, 1, 2)).astype(np.float32)
  got = S.stft_tf(torch.from_numpy(x).cuda(), 1024, 256).cpu().numpy()
  ref = O.stft_f32(x, 1024, 256)
  assert got.shape == ref.shape == (3, 20, 513, 2) and got.dtype == np.complex64
  assert _rel(got, ref) < TOL
  mag = S.magspec_tf(torch.from_numpy(x).cuda(), 1024, 256).cpu().numpy()
  assert _rel(mag, np.abs(ref)) < TOL


def test_r9y9_tf_noise_goldens():
  # reference tests/test_spectral.py:123-139 (places=3 on an f32 sum; see test_oracle_spectral)
  import torch
  from advoc_b200 import spectral as S
  np.random.seed(0)
  n1 = np.random.uniform(-1, 1, (1, 82432, 1, 1)).astype(np.float32)
  n2 = np.random.uniform(-1, 1, (1, 82432, 1, 2)).astype(np.float32)
  m2 = S.waveform_to_r9y9_melspec_tf(torch.from_numpy(n2).cuda()).cpu().numpy()
  m1 = S.waveform_to_r9y9_melspec_tf(torch.from_numpy(n1).cuda()).cpu().numpy()
  assert m2.shape == (1, 322, 80, 2) and m2.dtype == np.float32
  assert abs(float(m2[0, :, :, 0].astype(np.float64).sum()) - 18328.508) < 1e-2
  assert abs(float(m2[0, :, :, 1].astype(np.float64).sum()) - 18332.746) < 1e-2
  assert abs(float(m1[0, :, :, 0].astype(np.float64).sum()) - 18319.934) < 1e-2


def test_melspec_matches_oracle_config1(golden_dir):
  # BASELINE.json configs[0]: 1 x 22050 Hz 1 s mono through waveform_to_r9y9_melspec
  from advoc_b200 import spectral as S
  from oracle import spectral_np as O
  np.random.seed(0)
  x = np.random.uniform(-1, 1, 22050).astype(np.float32).reshape(-1, 1, 1)
  got = S.waveform_to_r9y9_melspec(x)
  ref = O.waveform_to_r9y9_melspec(x)
  assert got.shape == ref.shape == (87, 80, 1) and got.dtype == np.float64
  assert _rel(got, ref) < TOL
  xs = _sc09(golden_dir)
  got = S.waveform_to_r9y9_melspec(xs, fs=16000)
  ref = O.waveform_to_r9y9_melspec(xs, fs=16000)
  assert got.shape == ref.shape == (63, 80, 1)
  assert _rel(got, ref) < TOL


def test_tacotron2_preset_non_pow2_fft():
  # nfft=1200 (reference advoc/spectral.py:242-247) exercises the generic-length DFT path
  from advoc_b200 import spectral as S
  from oracle import spectral_np as O
  rng = np.random.RandomState(5)
  x = (rng.uniform(-1, 1, 24000) * np.hanning(24000)).astype(np.float32).reshape(-1, 1, 1)
  got = S.waveform_to_tacotron2_melspec(x)
  ref = O.waveform_to_tacotron2_melspec(x)
  assert got.shape == ref.shape == (80, 80, 1)
  assert _rel(got, ref) < TOL


def test_mel_matmuls_and_db_denorm(golden_dir):
  import torch
  from advoc_b200.model import SpectralUtil
  from oracle import spectral_np as O
  su = SpectralUtil()
  W = O.create_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80)
  Winv = O.create_inverse_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80)
  assert _rel(su.meltrans_np, W) < 1e-12 and _rel(su.invmeltrans_np, Winv) < 1e-9
  rng = np.random.RandomState(7)
  mag = np.abs(rng.randn(2, 37, 513, 1)).astype(np.float32)
  mel = su.mag_to_mel_linear_spec(torch.from_numpy(mag).cuda())
  ref_mel = O.mag_to_mel_linear_spec(mag, W)
  assert _rel(mel.cpu().numpy(), ref_mel) < TOL
  back = su.mel_linear_to_mag_spec(mel)
  assert _rel(back.cpu().numpy(), O.mel_linear_to_mag_spec(ref_mel.astype(np.float32), Winv)) < TOL
  # the r9y9 fixture as a realistic dB-normalised mel input
  r9 = np.load(os.path.join(golden_dir, 'mono_22k_r9y9_mel.npy')).T.copy()   # [325, 80]
  got = su.tacotron_mel_to_mag(r9)
  assert got.shape == (325, 513)
  assert _rel(got, O.tacotron_mel_to_mag(r9, Winv)) < TOL


def test_empty_and_error_contracts():
  import torch
  from advoc_b200 import spectral as S
  with pytest.raises(ValueError):
    S.stft(np.zeros((10, 2, 1), np.float32), 1024, 256)
  with pytest.raises(NotImplementedError):
    S.stft(np.zeros((10, 1, 2), np.float32), 1024, 256)
  with pytest.raises(ValueError):
    S.waveform_to_melspec(np.zeros((10, 1, 1), np.float64), 22050, 1024, 256)
  with pytest.raises(NotImplementedError):
    S.waveform_to_melspec_tf(torch.zeros(1, 10, 1, 1).cuda(), 22050, 1024, 256,
                             norm_allow_clipping=False)
  assert S.stft(np.zeros((0, 1, 1), np.float32), 1024, 256).shape == (0, 513, 1)
  # one partial frame, zero tail
  x = np.ones((100, 1, 1), np.float32)
  assert S.stft(x, 1024, 256).shape == (1, 513, 1)


def test_full_size_linearity_property():
  # BASELINE train-shape input [32, 65536(+768)] -> 259 frames: the STFT is linear
  import torch
  from advoc_b200 import spectral as S
  g = torch.Generator(device='cuda').manual_seed(0)
  a = torch.rand((32, 65536 + 768, 1, 1), device='cuda', generator=g) * 2 - 1
  b = torch.rand((32, 65536 + 768, 1, 1), device='cuda', generator=g) * 2 - 1
  Xa, Xb = S.stft_tf(a, 1024, 256), S.stft_tf(b, 1024, 256)
  Xab = S.stft_tf(a + 2 * b, 1024, 256)
  assert Xa.shape == (32, 259, 513, 1)
  err = (Xab - (Xa + 2 * Xb)).abs().pow(2).sum().sqrt() / Xab.abs().pow(2).sum().sqrt()
  assert float(err) < TOL
