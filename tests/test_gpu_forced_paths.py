"""The kernel-selection rule (csrc/conv_p2d.cu: conv_p2d_eligible) only picks the persistent patch
kernel for layers with several tiles per SM, which the small-batch parity tests never reach.  Re-run
the conv-stack and train-step parity suites in a child process with ADVOC_P2D_FORCE=1 so that every
eligible layer -- forward, input-gradient (TMA gate tiles, TMA reduce-add skip sums) -- goes through
it, and with ADVOC_NO_P2D=1 so that the per-tap kernel is covered for the same layers."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('var', ['ADVOC_P2D_FORCE', 'ADVOC_NO_P2D'])
def test_parity_suites_with_forced_kernel_choice(var):
  env = dict(os.environ)
  env[var] = '1'
  r = subprocess.run([sys.executable, '-m', 'pytest', '-x', '-q', '-m', 'gpu', 'tests/test_gpu_nets.py',
                      'tests/test_gpu_train.py', 'tests/test_gpu_melspecgan.py'],
                     cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
  assert r.returncode == 0, r.stdout[-3000:]
