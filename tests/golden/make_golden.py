"""Regenerates tests/golden/ from the reference checkout (run in the authoring container only;
/root/reference does not exist on the GPU box).

  sc09.wav               <- /root/reference/tests/audio/sc09.wav (byte copy, test data)
  mono.wav               <- /root/reference/tests/audio/mono.wav (byte copy, test data)
  mono_22k_r9y9_mel.npy  <- /root/reference/tests/audio/mono_22k_r9y9.pkl (pickle -> npy)
"""
import os
import pickle
import shutil

import numpy as np

REF = '/root/reference/tests/audio'
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == '__main__':
  shutil.copyfile(os.path.join(REF, 'sc09.wav'), os.path.join(HERE, 'sc09.wav'))
  shutil.copyfile(os.path.join(REF, 'mono.wav'), os.path.join(HERE, 'mono.wav'))
  with open(os.path.join(REF, 'mono_22k_r9y9.pkl'), 'rb') as f:
    mel = pickle.load(f, encoding='latin1')
  np.save(os.path.join(HERE, 'mono_22k_r9y9_mel.npy'), np.asarray(mel, dtype=np.float64))
  print('ok', mel.shape)
