"""Vocodes a directory of r9y9 mel spectrograms (.npy, [T, 80, 1] float64) into WAV files with
the adversarial vocoder on B200.

Drop-in for the reference's scripts/spectrogram_advoc.py (same flags).  Differences, all forced
by what exists here: `--model_ckpt` is a TF-1 checkpoint prefix as in the reference (`model.ckpt-N`, read
without TensorFlow by advoc_b200.tf_bundle) or an `.npz` written by advoc_b200.checkpoint (same variable
names); `--meta_fp` is accepted and ignored -- there is no TF graph; the reference's batch-1 chunk
loop (:80-95) is ONE batched forward, and the phase estimator is Griffin-Lim (--ngl, default 60)
because the LWS reconstruction of the third-party `lws` package is not restated.
"""
if __name__ == '__main__':
  from argparse import ArgumentParser
  import glob
  import os
  import sys

  import numpy as np

  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  from advoc_b200.audioio import save_as_wav
  from advoc_b200 import checkpoint, infer, spectral
  from advoc_b200.model import Advoc, AdvocSmall, Modes

  parser = ArgumentParser()
  parser.add_argument('--spec_dir', type=str, required=True, help='Directory of spectrograms')
  parser.add_argument('--out_dir', type=str, required=True, help='Directory for audio files')
  parser.add_argument('--model_ckpt', type=str, help='Adversarial vocoder checkpoint (TF prefix or .npz)')
  parser.add_argument('--meta_fp', type=str, help='Meta graph filepath (ignored)')
  parser.add_argument('--fs', type=int, help='Sample rate')
  parser.add_argument('--subseq_len', type=int, help='model subseq length')
  parser.add_argument('--ngl', type=int, help='Griffin-Lim iterations')
  parser.set_defaults(spec_dir=None, out_dir=None, model_ckpt=None, meta_fp=None, fs=22050,
                      subseq_len=256, ngl=60)
  args = parser.parse_args()

  if not os.path.isdir(args.out_dir):
    os.makedirs(args.out_dir)

  model = None
  if args.model_ckpt is None:
    print('Warning: Model checkpoint not specified, using pseudoinverse+Griffin-Lim heuristic to vocode')
  else:
    params, step = checkpoint.load_params(args.model_ckpt)
    cls = AdvocSmall if checkpoint.infer_model_type(params) == 'small' else Advoc
    model = cls(Modes.INFER, params=params)
    model.audio_fs = args.fs
    model.subseq_len = args.subseq_len
    print('Restored %s (step %d)' % (args.model_ckpt, step))

  for spec_fp in sorted(glob.glob(os.path.join(args.spec_dir, '*.npy'))):
    spec_fn = os.path.splitext(os.path.split(spec_fp)[1])[0]
    wave_fp = os.path.join(args.out_dir, spec_fn + '.wav')
    spec = np.load(spec_fp)
    if model is None:
      wave = spectral.r9y9_melspec_to_waveform(spec.astype(np.float64), fs=args.fs,
                                               phase_estimation='gl%d' % args.ngl)
    else:
      gen_mag = infer.mel_to_mag(model, spec[:, :, 0].astype(np.float32), input_kind='dbnorm')
      wave = spectral.magspec_to_waveform_griffin_lim(gen_mag[:, :, np.newaxis], 1024, 256,
                                                      ngl=args.ngl)
    save_as_wav(wave_fp, args.fs, wave)
    print(wave_fp)
