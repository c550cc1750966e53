"""Parameter checkpoints keyed by the reference's TF variable names.

The reference checkpoints through `tf.train.MonitoredTrainingSession` / `Saver`
(models/advoc/train_evaluate.py:61-64, scripts/spectrogram_advoc.py:57-64); the variable names
(`generator/encoder_N/conv2d/{kernel,bias}`, `generator/decoder_N/conv2d_transpose/{kernel,bias}`,
`discriminator/layer_N/conv2d/{kernel,bias}`) and layouts (HWIO / HWOI) are kept, so a converted TF
checkpoint (any name->array dump) loads directly.  Container: a flat `.npz`.  Reading the TF
bundle format itself (.index SSTable + .data) is not implemented.
"""
import numpy as np
import torch


def save_params(path, params, step=0, extra=None):
  arrays = {k: v.detach().cpu().numpy() for k, v in params.items()}
  arrays['global_step'] = np.asarray(step, dtype=np.int64)
  for k, v in (extra or {}).items():
    arrays['extra/' + k] = np.asarray(v)
  np.savez(path, **arrays)


def load_params(path, device='cuda'):
  """-> (params dict of CUDA float32 tensors, global_step)."""
  z = np.load(path)
  step = int(z['global_step']) if 'global_step' in z.files else 0
  P = {k: torch.from_numpy(np.ascontiguousarray(z[k], dtype=np.float32)).to(device)
       for k in z.files if k.startswith(('generator/', 'discriminator/'))}
  return P, step


def infer_model_type(params):
  """'small' (ngf 32, 5+5 layers) or 'regular' (ngf 64, 8+8) from the parameter shapes."""
  ngf = params['generator/encoder_1/conv2d/kernel'].shape[3]
  n_enc = len([k for k in params if k.startswith('generator/encoder_') and k.endswith('/kernel')])
  if ngf == 32 and n_enc == 5:
    return 'small'
  if ngf == 64 and n_enc == 8:
    return 'regular'
  raise ValueError('unrecognised AdVoc variant: ngf=%d, %d encoders' % (ngf, n_enc))
