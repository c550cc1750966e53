"""Directory of waveforms -> directory of r9y9 mel spectrograms (.npy [T, 80, 1] float64).
Drop-in for the reference's scripts/audio_to_spectrogram.py (same flags); the feature kernel is
the fused STFT->mel->dB kernel of libadvoc_b200.so."""
if __name__ == '__main__':
  from argparse import ArgumentParser
  import glob
  import os
  import sys

  import numpy as np

  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  from advoc_b200.audioio import decode_audio
  from advoc_b200.spectral import waveform_to_r9y9_melspec

  parser = ArgumentParser()
  parser.add_argument('--wave_dir', type=str, required=True, help='Directory of audio files')
  parser.add_argument('--out_dir', type=str, required=True, help='Directory for spectrograms')
  parser.add_argument('--fs', type=int, help='Sample rate')
  parser.add_argument('--data_fast_wav', action='store_true', dest='data_fast_wav',
                      help='If set, provides faster loading of standard WAV files via scipy')
  parser.set_defaults(wave_dir=None, out_dir=None, fs=22050, data_fast_wav=False)
  args = parser.parse_args()

  if not os.path.isdir(args.out_dir):
    os.makedirs(args.out_dir)
  for wave_fp in sorted(glob.glob(os.path.join(args.wave_dir, '*'))):
    wave_fn = os.path.splitext(os.path.split(wave_fp)[1])[0]
    fs, wave = decode_audio(wave_fp, fs=args.fs if not args.data_fast_wav else None,
                            fastwav=args.data_fast_wav, mono=True, normalize=True)
    spec = waveform_to_r9y9_melspec(wave.astype(np.float32), fs=fs)
    np.save(os.path.join(args.out_dir, wave_fn + '.npy'), spec)
