"""On-GPU extract stage of the reference's input pipeline.

The reference computes features per file inside tf.data (`stft_tf` -> `tf.abs`, then
`tf.contrib.signal.frame` slicing of the FEATURE sequence; advoc/loader.py:117-128,133-186).
Here one kernel (`advoc_stft_f32`, magnitude-only output) produces the whole-file magnitude
spectrogram on the GPU and the slices are strided views of it; file decoding stays on the host
(advoc_b200.audioio).  The datacfg semantics are kept: slice_len frames per example, hop =
round(slice_len * (1 - overlap_ratio)), optional random start offset, zero padding of the last
partial slice (slice_pad_end), first-slice-only mode.
"""
import numpy as np
import torch

from advoc_b200 import spectral


def extract_magspec_slices(wav, slice_len=256, nfft=1024, nhop=256, slice_overlap_ratio=0.,
                           slice_randomize_offset=False, slice_pad_end=True, slice_first_only=False,
                           rng=None):
  """wav: float32 [nsamps, 1, 1] (numpy or cuda tensor) -> cuda float32 [n_slices, slice_len, bins, 1].

  reference: advoc/loader.py:117-128 (extract_type='magspec') and :133-186 (`_parallel_slice`)."""
  if slice_overlap_ratio < 0:
    raise ValueError('Slice overlap must be nonnegative')
  slice_hop = int(round(slice_len * (1. - slice_overlap_ratio)))
  if slice_hop < 1:
    raise ValueError('Overlap ratio too high')
  if isinstance(wav, np.ndarray):
    wav = torch.from_numpy(np.ascontiguousarray(wav, dtype=np.float32)).cuda()
  if wav.dim() != 3 or wav.shape[1] != 1 or wav.shape[2] != 1:
    raise ValueError()
  feats = spectral.magspec_tf(wav.reshape(1, -1, 1, 1), nfft, nhop)[0]      # [frames, bins, 1]
  if slice_randomize_offset:
    rng = np.random if rng is None else rng
    feats = feats[int(rng.randint(slice_len)):]
  n = feats.shape[0]
  if slice_pad_end:
    n_slices = -(-n // slice_hop) if n > 0 else 0
    need = (n_slices - 1) * slice_hop + slice_len if n_slices > 0 else 0
    if need > n:
      feats = torch.cat([feats, feats.new_zeros((need - n,) + tuple(feats.shape[1:]))], 0)
  else:
    n_slices = (n - slice_len) // slice_hop + 1 if n >= slice_len else 0
  if n_slices == 0:
    return feats.new_zeros((0, slice_len) + tuple(feats.shape[1:]))
  slices = feats.unfold(0, slice_len, slice_hop).permute(0, 3, 1, 2)        # [n, slice_len, bins, 1]
  if slice_first_only:
    slices = slices[:1]
  return slices.contiguous()


def batches(slices, batch_size, drop_remainder=True):
  """Yields [batch, slice_len, bins, 1] blocks (the reference batches with drop_remainder,
  advoc/loader.py:195-199)."""
  n = slices.shape[0]
  stop = n - n % batch_size if drop_remainder else n
  for i in range(0, stop, batch_size):
    yield slices[i:i + batch_size]
