"""Creates a directory of mel spectrograms (.npy, [64, 80, 1] float32 in [0, 1]) with a trained MelspecGAN
generator on B200.

Drop-in for the reference's scripts/generate_spectrogram.py (same flags).  `--ckpt_fp` is a TF-1
checkpoint prefix as in the reference (read without TensorFlow by advoc_b200.tf_bundle) or an `.npz`
with the same variable names; `--meta_fp` is accepted and ignored -- the inference graph of
models/melspecgan/infer.py (z -> G(z, training=False) -> feats_denorm) is built into
advoc_b200.melspecgan instead of being imported from a meta graph.
"""
if __name__ == '__main__':
  from argparse import ArgumentParser
  import os
  import sys

  import numpy as np
  import torch

  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  from advoc_b200 import checkpoint
  from advoc_b200.melspecgan import MelspecGANGenerator, Z_DIM, init_params

  parser = ArgumentParser()
  parser.add_argument('--out_dir', type=str, required=True, help='Directory for spectrograms')
  parser.add_argument('--ckpt_fp', type=str, help='MelspecGAN checkpoint (TF prefix or .npz)')
  parser.add_argument('--meta_fp', type=str, help='Meta graph filepath (ignored)')
  parser.add_argument('--n', type=int, help='Total number of spectrograms to generate')
  parser.add_argument('--b', type=int, help='Number of spectrograms to generate per batch')
  parser.add_argument('--seed', type=int, help='Seed of the latent draws')
  parser.set_defaults(out_dir=None, ckpt_fp=None, meta_fp=None, n=1000, b=10, seed=0)
  args = parser.parse_args()

  if not os.path.isdir(args.out_dir):
    os.makedirs(args.out_dir)

  if args.ckpt_fp is None:
    print('Warning: no checkpoint given, generating from randomly initialised weights')
    params, moving = init_params(), None
  else:
    params, step = checkpoint.load_params(args.ckpt_fp)
    params = {k: v for k, v in params.items() if k.startswith('G/')}
    moving = checkpoint.load_bn_moving(args.ckpt_fp) or None
    print('Restored from step {}'.format(step))
  dim = params['G/upconv_4/W'].shape[3]
  if not any(k.startswith('D/') for k in params):     # the engine wants both nets; D is never run here
    params.update({k: v for k, v in init_params(dim=dim).items() if k.startswith('D/')})
  G = MelspecGANGenerator(dim=dim, params=params, moving=moving)

  gen = torch.Generator(device='cuda')
  gen.manual_seed(args.seed)
  for i in range(0, args.n, args.b):
    z = torch.randn((args.b, Z_DIM), dtype=torch.float32, device='cuda', generator=gen)   # samp_z
    G_z = (G(z, training=False) + 1.) * 0.5                                                # feats_denorm
    for j, s in enumerate(G_z.cpu().numpy()):
      np.save(os.path.join(args.out_dir, '{}.npy'.format(str(j + i).zfill(9))), s.astype(np.float32))
