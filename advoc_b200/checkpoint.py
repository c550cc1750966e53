"""Parameter checkpoints keyed by the reference's TF variable names.

The reference checkpoints through `tf.train.MonitoredTrainingSession` / `Saver`
(models/advoc/train_evaluate.py:61-64, scripts/spectrogram_advoc.py:57-64); the variable names
(`generator/encoder_N/conv2d/{kernel,bias}`, `generator/decoder_N/conv2d_transpose/{kernel,bias}`,
`discriminator/layer_N/conv2d/{kernel,bias}`) and layouts (HWIO / HWOI) are kept, so a converted TF
checkpoint loads directly.  Containers: a flat `.npz` (written here), or a TF-1 checkpoint prefix
(`model.ckpt-N` with its `.index` / `.data-*` files, or a train directory holding a `checkpoint`
state file), parsed by `advoc_b200.tf_bundle` without TensorFlow.
"""
import os

import numpy as np
import torch

from advoc_b200 import tf_bundle

# variable scopes of the reference graphs: AdVoc (advoc_model.py:247-248), MelspecGAN (train.py:117-118)
_SCOPES = ('generator/', 'discriminator/', 'G/', 'D/')
_SLOT_SUFFIXES = ('/Adam', '/Adam_1')   # tf.train.AdamOptimizer slot variables m, v
_NON_TRAINABLE = ('/moving_mean', '/moving_variance')   # batch-norm inference statistics (MelspecGAN)


def save_params(path, params, step=0, extra=None):
  """Save {name: tensor} + global_step.  `path` ending in `.npz`: one flat npz.  Anything else is taken as a
  TF-1 checkpoint prefix (e.g. `train_dir/model.ckpt-1000`): `.index` + `.data-00000-of-00001` in the
  layout `tf.train.Saver` restores, plus the directory's `checkpoint` state file (so that
  `tf.train.latest_checkpoint` / `resolve_tf_prefix(train_dir)` find it).  `extra`: further arrays, stored
  under `extra/<key>` (npz) or under their own names (TF prefix: Adam slots, moving statistics ...)."""
  arrays = {k: v.detach().cpu().numpy() for k, v in params.items()}
  arrays['global_step'] = np.asarray(step, dtype=np.int64)
  if path.endswith('.npz'):
    for k, v in (extra or {}).items():
      arrays['extra/' + k] = np.asarray(v)
    np.savez(path, **arrays)
    return
  for k, v in (extra or {}).items():
    arrays[k] = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
  tf_bundle.write_bundle(path, arrays)
  d, base = os.path.split(path)
  with open(os.path.join(d, 'checkpoint'), 'w') as f:
    f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


def resolve_tf_prefix(path):
  """`path` as a TF checkpoint prefix, or None: accepts the prefix itself, one of its files, or a
  train directory with a `checkpoint` state file (tf.train.latest_checkpoint)."""
  if os.path.isdir(path):
    return tf_bundle.latest_checkpoint(path)
  if tf_bundle.is_bundle(path):
    return path
  for suffix in ('.index', '.meta'):
    if path.endswith(suffix) and tf_bundle.is_bundle(path[:-len(suffix)]):
      return path[:-len(suffix)]
  stem = path.rsplit('.data-', 1)[0]
  if stem != path and tf_bundle.is_bundle(stem):
    return stem
  return None


def load_arrays(path):
  """-> {name: numpy array} of every variable stored at `path` (.npz or TF checkpoint)."""
  prefix = resolve_tf_prefix(path)
  if prefix is not None:
    return tf_bundle.read_bundle(prefix)
  if not os.path.isfile(path):
    raise FileNotFoundError('no .npz file or TensorFlow checkpoint at %s' % path)
  z = np.load(path)
  return {k: z[k] for k in z.files}


def load_params(path, device='cuda'):
  """-> (params dict of float32 tensors on `device`, global_step).  Model variables only: the Adam
  slots of a TF training checkpoint (`<var>/Adam`, `<var>/Adam_1`, `beta*_power`) are skipped here,
  see `load_adam_slots`, and so are batch-norm moving statistics (the engines here run the
  training-mode graph, batch statistics)."""
  A = load_arrays(path)
  step = int(A['global_step']) if 'global_step' in A else 0
  P = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(device)
       for k, v in A.items() if k.startswith(_SCOPES) and not k.endswith(_SLOT_SUFFIXES + _NON_TRAINABLE)}
  return P, step


def load_bn_moving(path, device='cuda'):
  """-> {name: tensor} of the batch-norm moving statistics in the checkpoint (MelspecGAN's inference
  graph restores them with the generator variables: models/melspecgan/infer.py:22-24)."""
  A = load_arrays(path)
  return {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(device)
          for k, v in A.items() if k.startswith(_SCOPES) and k.endswith(_NON_TRAINABLE)}


def load_adam_slots(path, device='cuda'):
  """-> (m, v, beta_powers): the first / second moment slots of a TF training checkpoint keyed by
  their variable's name, and {'beta1_power': .., 'beta2_power': ..} when stored (the step count of
  tf.train.AdamOptimizer lives in those powers)."""
  A = load_arrays(path)
  m, v, powers = {}, {}, {}
  for k, a in A.items():
    t = lambda: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
    if k.startswith(_SCOPES) and k.endswith('/Adam'):
      m[k[:-len('/Adam')]] = t()
    elif k.startswith(_SCOPES) and k.endswith('/Adam_1'):
      v[k[:-len('/Adam_1')]] = t()
    elif k.split('/')[-1].startswith(('beta1_power', 'beta2_power')):
      powers[k] = float(a)
  return m, v, powers


def infer_model_type(params):
  """'small' (ngf 32, 5+5 layers) or 'regular' (ngf 64, 8+8) from the parameter shapes."""
  ngf = params['generator/encoder_1/conv2d/kernel'].shape[3]
  n_enc = len([k for k in params if k.startswith('generator/encoder_') and k.endswith('/kernel')])
  if ngf == 32 and n_enc == 5:
    return 'small'
  if ngf == 64 and n_enc == 8:
    return 'regular'
  raise ValueError('unrecognised AdVoc variant: ngf=%d, %d encoders' % (ngf, n_enc))
