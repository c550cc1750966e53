"""advoc_b200.audioio against the reference's own test expectations (tests/test_audioio.py,
scipy path; the librosa/mp3 cases need librosa, absent here).  CPU only."""
import os
import tempfile

import numpy as np
import pytest

from advoc_b200.audioio import decode_audio, save_as_wav


def test_scipy_decode_audio(golden_dir):
  # reference tests/test_audioio.py:19-40
  wav = os.path.join(golden_dir, 'mono.wav')
  fs, x = decode_audio(wav, fastwav=True)
  assert x.dtype == np.float32 and fs == 44100 and x.shape == (164864, 1, 1)
  assert abs(float(x.min()) - -0.474823) < 1e-6 and abs(float(x.max()) - 0.397278) < 1e-6
  with pytest.raises(ValueError):
    decode_audio(wav, fs=22050, fastwav=True)
  fs, x = decode_audio(wav, normalize=True, fastwav=True)
  assert abs(float(np.abs(x).max()) - 1.) < 1e-8
  with pytest.raises(ValueError):
    decode_audio(os.path.join(golden_dir, 'mono_22k_r9y9_mel.npy'), fastwav=True)   # not a WAV


def test_stereo_and_mono_mix(tmp_path):
  from scipy.io import wavfile
  rng = np.random.RandomState(0)
  st = (rng.uniform(-0.5, 0.5, (1000, 2)) * 32767).astype(np.int16)
  fp = str(tmp_path / 'stereo.wav')
  wavfile.write(fp, 44100, st)
  fs, x = decode_audio(fp, fastwav=True)
  assert x.shape == (1000, 1, 2)
  fs, m = decode_audio(fp, mono=True, fastwav=True)
  assert m.shape == (1000, 1, 1)
  assert np.allclose(m[:, 0, 0], x[:, 0, :].mean(axis=1))


def test_save_as_wav_roundtrip(golden_dir):
  # reference tests/test_audioio.py:80-96
  fs, x = decode_audio(os.path.join(golden_dir, 'mono.wav'), fastwav=True)
  with tempfile.NamedTemporaryFile(suffix='.wav') as tf:
    with pytest.raises(ValueError):
      save_as_wav(tf.name, fs, x[:, 0])
    with pytest.raises(ValueError):
      save_as_wav(tf.name, fs, np.concatenate([x, x], axis=1))
    with pytest.raises(NotImplementedError):
      save_as_wav(tf.name, fs, np.concatenate([x, x], axis=2))
    save_as_wav(tf.name, fs, x)
    fs2, x2 = decode_audio(tf.name, fastwav=True)
    assert fs2 == fs and np.array_equal(x, x2)
