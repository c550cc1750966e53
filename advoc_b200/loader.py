"""On-GPU extract stage of the reference's input pipeline.

The reference computes features per file inside tf.data (`stft_tf` -> `tf.abs` or
`waveform_to_melspec_tf`, then `tf.contrib.signal.frame` slicing of the FEATURE sequence and of the
AUDIO in parallel; advoc/loader.py:98-186).  Here one kernel (`advoc_stft_f32` with a magnitude-only
output, or the fused `advoc_melspec_f32`) produces the whole-file features on the GPU and the slices
are strided views of it; file decoding stays on the host (advoc_b200.audioio).  The datacfg semantics
are kept: slice_len frames per example, hop = round(slice_len * (1 - overlap_ratio)), optional
random start offset (the audio offset follows the feature offset), zero padding of the last partial
slice (slice_pad_end), first-slice-only mode, batches with drop_remainder.
"""
import numpy as np
import torch

from advoc_b200 import spectral


def _frame(t, frame_len, frame_hop, pad_end):
  """tf.contrib.signal.frame(t, frame_len, frame_hop, pad_end, pad_value=0, axis=0):
  pad_end -> ceil(n / hop) frames (the tail zero-padded), else floor((n - len) / hop) + 1."""
  n = t.shape[0]
  if pad_end:
    n_frames = -(-n // frame_hop) if n > 0 else 0
    need = (n_frames - 1) * frame_hop + frame_len if n_frames > 0 else 0
    if need > n:
      t = torch.cat([t, t.new_zeros((need - n,) + tuple(t.shape[1:]))], 0)
  else:
    n_frames = (n - frame_len) // frame_hop + 1 if n >= frame_len else 0
  if n_frames == 0:
    return t.new_zeros((0, frame_len) + tuple(t.shape[1:]))
  # unfold puts the window last: [n_frames, rest..., frame_len] -> [n_frames, frame_len, rest...]
  f = t.unfold(0, frame_len, frame_hop)
  return f.permute(0, f.dim() - 1, *range(1, f.dim() - 1)).contiguous()


def extract_features(wav, extract_type, audio_fs=22050, nfft=1024, nhop=256):
  """wav f32 [nsamps, 1, nch] on the GPU -> features [ntsteps, nfeats, nch] (advoc/loader.py:98-128):
  None -> the audio itself, 'magspec' -> |stft_tf|, 'melspec' -> waveform_to_melspec_tf."""
  if extract_type is None:
    return wav
  if extract_type == 'magspec':
    return spectral.magspec_tf(wav[None], nfft, nhop)[0]
  if extract_type == 'melspec':
    return spectral.waveform_to_melspec_tf(wav[None], fs=audio_fs, nfft=nfft, nhop=nhop)[0]
  raise ValueError()


def parallel_slice(features, audio, slice_len, audio_fs, feature_fs, slice_overlap_ratio=0.,
                   slice_randomize_offset=False, slice_pad_end=False, slice_first_only=False, rng=None):
  """`_parallel_slice` of advoc/loader.py:133-186: paired feature slices [n, slice_len, nfeats, nch] and
  audio slices [n, slice_len * samples_per_step, 1, nch]."""
  if slice_overlap_ratio < 0:
    raise ValueError('Slice overlap must be nonnegative')
  slice_hop = int(round(slice_len * (1. - slice_overlap_ratio)))
  if slice_hop < 1:
    raise ValueError('Overlap ratio too high')
  nsamps_per_tstep = float(audio_fs) / float(feature_fs)
  audio_slice_len = int(round(slice_len * nsamps_per_tstep) + 1e-4)
  audio_slice_hop = int(round(slice_hop * nsamps_per_tstep) + 1e-4)
  if slice_randomize_offset:
    rng = np.random if rng is None else rng
    start = int(rng.randint(slice_len))
    start_audio = int(np.round(np.float32(start) * np.float32(nsamps_per_tstep) + 1e-4))
    audio = audio[start_audio:]
    features = features[start:]
  feature_slices = _frame(features, slice_len, slice_hop, slice_pad_end)
  audio_slices = _frame(audio, audio_slice_len, audio_slice_hop, slice_pad_end)
  if slice_first_only:
    feature_slices, audio_slices = feature_slices[:1], audio_slices[:1]
  return feature_slices, audio_slices


def extract_magspec_slices(wav, slice_len=256, nfft=1024, nhop=256, slice_overlap_ratio=0.,
                           slice_randomize_offset=False, slice_pad_end=True, slice_first_only=False,
                           rng=None):
  """wav: float32 [nsamps, 1, 1] (numpy or cuda tensor) -> cuda float32 [n_slices, slice_len, bins, 1].

  reference: advoc/loader.py:117-128 (extract_type='magspec') and :133-186 (`_parallel_slice`)."""
  return extract_slices(wav, 'magspec', slice_len, nfft=nfft, nhop=nhop, slice_overlap_ratio=slice_overlap_ratio,
                        slice_randomize_offset=slice_randomize_offset, slice_pad_end=slice_pad_end,
                        slice_first_only=slice_first_only, rng=rng)[0]


def extract_slices(wav, extract_type, slice_len, audio_fs=22050, nfft=1024, nhop=256, **slice_kw):
  """One decoded file -> (feature_slices, audio_slices) on the GPU.  extract_type 'melspec' is what
  models/melspecgan/train.py:20-42 trains on, 'magspec' what models/advoc/train_evaluate.py does."""
  if slice_kw.get('slice_overlap_ratio', 0.) < 0:
    raise ValueError('Slice overlap must be nonnegative')
  if int(round(slice_len * (1. - slice_kw.get('slice_overlap_ratio', 0.)))) < 1:
    raise ValueError('Overlap ratio too high')
  if isinstance(wav, np.ndarray):
    wav = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32)).cuda()
  if wav.dim() != 3 or wav.shape[1] != 1:
    raise ValueError()
  feats = extract_features(wav, extract_type, audio_fs, nfft, nhop)
  feature_fs = audio_fs if extract_type is None else audio_fs / nhop
  return parallel_slice(feats, wav, slice_len, audio_fs, feature_fs, **slice_kw)


def batches(slices, batch_size, drop_remainder=True):
  """Yields [batch, slice_len, bins, 1] blocks (the reference batches with drop_remainder,
  advoc/loader.py:195-199)."""
  n = slices.shape[0]
  stop = n - n % batch_size if drop_remainder else n
  for i in range(0, stop, batch_size):
    yield slices[i:i + batch_size]


def decode_extract_and_batch(fps, batch_size, slice_len, audio_fs=22050, audio_mono=True, audio_normalize=False,
                             decode_fastwav=False, decode_parallel_calls=1, extract_type=None, extract_nfft=1024,
                             extract_nhop=256, extract_parallel_calls=1, repeat=False, shuffle=False,
                             shuffle_buffer_size=None, slice_first_only=False, slice_randomize_offset=False,
                             slice_overlap_ratio=0, slice_pad_end=False, prefetch_size=None, prefetch_gpu_num=None,
                             seed=None):
  """The reference's entry point with the reference's signature (advoc/loader.py:8-218), as a Python
  generator of (x_feats [b, slice_len, nfeats, nch], x_audio [b, slice_len * nhop, 1, nch]) CUDA
  tensors instead of a pair of graph tensors.  Files are decoded on the host (audioio.decode_audio),
  everything after that runs on the GPU; the *_parallel_calls / prefetch_* knobs of tf.data have no
  counterpart and are ignored.  shuffle: file order every epoch plus a slice buffer of
  `shuffle_buffer_size` examples, like the reference's two `dataset.shuffle` calls."""
  from advoc_b200 import audioio
  if extract_type not in (None, 'magspec', 'melspec'):
    raise ValueError()
  rng = np.random.RandomState(seed)
  fps = list(fps)
  dev = torch.device('cuda', prefetch_gpu_num if (prefetch_gpu_num is not None and prefetch_gpu_num >= 0)
                     else torch.cuda.current_device())
  buf_f, buf_a = [], []
  cap = max(int(shuffle_buffer_size or 1), batch_size) if shuffle else batch_size

  def pop_batch():
    idx = rng.permutation(len(buf_f))[:batch_size] if shuffle else np.arange(batch_size)
    f = torch.stack([buf_f[i] for i in idx])
    a = torch.stack([buf_a[i] for i in idx])
    for i in sorted(idx, reverse=True):
      del buf_f[i], buf_a[i]
    return f, a

  while True:
    order = rng.permutation(len(fps)) if shuffle else np.arange(len(fps))
    for fi in order:
      _, wav = audioio.decode_audio(fps[fi], fs=audio_fs, mono=audio_mono, normalize=audio_normalize,
                                    fastwav=decode_fastwav)
      wav = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32)).to(dev)
      fs_, as_ = extract_slices(wav, extract_type, slice_len, audio_fs, extract_nfft, extract_nhop,
                                slice_overlap_ratio=slice_overlap_ratio, slice_randomize_offset=slice_randomize_offset,
                                slice_pad_end=slice_pad_end, slice_first_only=slice_first_only, rng=rng)
      n = min(fs_.shape[0], as_.shape[0])
      buf_f.extend(fs_[:n].unbind(0))
      buf_a.extend(as_[:n].unbind(0))
      while len(buf_f) >= cap and len(buf_f) >= batch_size:
        yield pop_batch()
    if not repeat:
      break
  while len(buf_f) >= batch_size:       # drop_remainder=True
    yield pop_batch()
