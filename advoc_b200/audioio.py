"""Drop-in for the reference's `advoc.audioio` (host file I/O; no kernels involved).

reference: advoc/audioio.py:9-68 (`decode_audio`), :71-93 (`save_as_wav`).  Same shapes
([nsamps, 1, nch] float32), same exceptions.  librosa (mp3 decode / resampling) is optional like
in the reference; without it only the scipy fast path and plain WAV files are available.
"""
import numpy as np
from scipy.io import wavfile

try:
  import librosa  # noqa: F401
except ImportError:
  librosa = None


def decode_audio(fp, fs=None, mono=False, normalize=False, fastwav=False):
  """Decodes an audio file into (fs, float32 [nsamps, 1, nch]).  advoc/audioio.py:9-68."""
  if fastwav:
    try:
      orig_fs, x = wavfile.read(fp)
    except Exception:
      raise ValueError('Error encountered when decoding WAV file.')
    if fs is not None and fs != orig_fs:
      raise ValueError('Fastwav cannot resample audio.')
    fs = orig_fs
    if x.dtype == np.int16:
      x = x.astype(np.float32) / np.float32(32768.)
    elif x.dtype == np.float32:
      pass
    else:
      raise ValueError('Fastwav cannot process atypical WAV files.')
  else:
    if librosa is None:
      raise Exception('Please install librosa')
    try:
      x, fs = librosa.core.load(fp, sr=fs, mono=False)
    except Exception:
      raise ValueError('Error encountered when decoding audio file.')
    if x.ndim == 2:
      x = np.swapaxes(x, 0, 1)
  assert x.dtype == np.float32
  if x.ndim == 1:
    nsamps, nch = x.shape[0], 1
  else:
    nsamps, nch = x.shape
  x = np.reshape(x, [nsamps, 1, nch])
  if mono:
    x = np.mean(x, 2, keepdims=True)
  if normalize:
    factor = np.max(np.abs(x))
    if factor > 0:
      x = x / factor
  return fs, x


def save_as_wav(fp, fs, x):
  """Saves float32 [?, 1, 1] as signed 16-bit PCM WAV.  advoc/audioio.py:71-93."""
  try:
    nsamps, nfeats, nch = x.shape
  except ValueError:
    raise ValueError('Incorrect number of input dimesions.')
  if nfeats != 1:
    raise ValueError('Incorrect input dimesions.')
  if nch != 1:
    raise NotImplementedError('Can only save monaural WAV for now.')
  y = np.clip(np.copy(x[:, 0, 0]) * 32768., -32768., 32767.).astype(np.int16)
  wavfile.write(fp, fs, y)
