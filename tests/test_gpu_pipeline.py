"""GPU tests of the pieces around the nets: loader extract stage, batched inference with the
reference's pad/trim rule, checkpoints, the CLI entry points, and data-parallel equivalence."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
  a = torch.as_tensor(a).detach().double().cpu()
  b = torch.as_tensor(b).detach().double().cpu()
  return float((a - b).norm() / b.norm().clamp_min(1e-300))


def test_loader_extract_slices_follow_reference_framing():
  # advoc/loader.py:133-186 semantics on the magnitude STFT of a whole file
  from advoc_b200 import loader
  from oracle import spectral_np as O
  rng = np.random.RandomState(0)
  wav = rng.uniform(-1, 1, (100000, 1, 1)).astype(np.float32)
  mag = O.magspec_f32(wav[np.newaxis], 1024, 256)[0]                 # [391, 513, 1]
  s = loader.extract_magspec_slices(wav, slice_len=256, slice_overlap_ratio=0.25)
  hop = 192
  n = -(-391 // hop)
  assert s.shape == (n, 256, 513, 1)
  ref = np.zeros(((n - 1) * hop + 256, 513, 1), np.float32)
  ref[:391] = mag
  for i in range(n):
    assert _rel(s[i], ref[i * hop:i * hop + 256]) < 1e-4
  s2 = loader.extract_magspec_slices(wav, slice_len=256, slice_overlap_ratio=0., slice_pad_end=False)
  assert s2.shape == (1, 256, 513, 1)
  assert loader.extract_magspec_slices(wav, slice_first_only=True).shape[0] == 1
  with pytest.raises(ValueError):
    loader.extract_magspec_slices(wav, slice_overlap_ratio=1.0)
  assert len(list(loader.batches(s, 2))) == n // 2


def test_batched_inference_pad_and_trim_rule(golden_dir):
  # scripts/spectrogram_advoc.py:80-95: pad to floor(T/256)*256 + 256 frames in the MAGNITUDE
  # domain, run 256-frame chunks, concatenate, trim -- here as one batched forward
  from advoc_b200 import infer
  from advoc_b200.model import AdvocSmall, Modes
  from oracle import nets_torch as O
  from oracle import spectral_np as OS
  P = O.init_params(O.SMALL, seed=0)
  model = AdvocSmall(Modes.INFER, params={k: v.cuda() for k, v in P.items()})
  mel = np.load(os.path.join(golden_dir, 'mono_22k_r9y9_mel.npy')).T.copy()        # [325, 80]
  got = infer.mel_to_mag(model, mel.astype(np.float32), input_kind='dbnorm', dropout=None)
  assert got.shape == (325, 513)
  Winv = OS.create_inverse_mel_filterbank(22050, 1024, fmin=125., fmax=7600., n_mels=80)
  X = OS.tacotron_mel_to_mag(mel, Winv).astype(np.float32)
  X = np.pad(X, ([0, 512 - 325], [0, 0]), 'constant').reshape(2, 256, 513, 1)
  ref = O.generator(P, torch.from_numpy(X), O.SMALL).reshape(512, 513)[:325]
  assert _rel(got, ref) < 1e-3


def test_checkpoint_roundtrip_and_cli(tmp_path, golden_dir):
  from advoc_b200 import checkpoint, nets
  P = nets.init_params(32, 32, 5, seed=3)
  ck = str(tmp_path / 'model.npz')
  checkpoint.save_params(ck, P, step=123)
  Q, step = checkpoint.load_params(ck)
  assert step == 123 and sorted(Q) == sorted(P)
  assert all(torch.equal(P[k], Q[k]) for k in P)
  assert checkpoint.infer_model_type(Q) == 'small'
  # the inference entry point, with and without a checkpoint
  spec_dir, out_dir = tmp_path / 'specs', tmp_path / 'wavs'
  spec_dir.mkdir()
  mel = np.load(os.path.join(golden_dir, 'mono_22k_r9y9_mel.npy')).T[:100, :, np.newaxis]
  np.save(str(spec_dir / 'utt.npy'), mel)
  for extra in ([], ['--model_ckpt', ck, '--meta_fp', 'ignored.meta']):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'spectrogram_advoc.py'),
                        '--spec_dir', str(spec_dir), '--out_dir', str(out_dir), '--ngl', '3'] + extra,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    assert r.returncode == 0, r.stdout
    from scipy.io import wavfile
    fs, w = wavfile.read(str(out_dir / 'utt.wav'))
    assert fs == 22050 and w.dtype == np.int16 and w.shape == (99 * 256 + 1024,)
  # wav -> mel CLI
  wdir, sdir = tmp_path / 'w', tmp_path / 's'
  wdir.mkdir()
  import shutil
  shutil.copyfile(os.path.join(golden_dir, 'sc09.wav'), str(wdir / 'a.wav'))
  r = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'audio_to_spectrogram.py'),
                      '--wave_dir', str(wdir), '--out_dir', str(sdir), '--data_fast_wav'],
                     stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
  assert r.returncode == 0, r.stdout
  m = np.load(str(sdir / 'a.npy'))
  assert m.shape == (63, 80, 1) and m.dtype == np.float64 and 0 <= m.min() and m.max() <= 1


def test_data_parallel_gradient_equivalence():
  """the mean of per-rank gradients over equal shards == the gradient of the
  global batch.  Emulated on one GPU: two engines on the two halves (summed, x 1/2) vs one engine
  on the whole batch, exact-fp32 math, injected dropout masks."""
  from advoc_b200 import _native as N
  from advoc_b200 import nets
  from advoc_b200.train import TrainEngine
  P = nets.init_params(32, 32, 5, seed=0)
  spec = nets.GenSpec(32, 5, (5, 4))
  g = torch.Generator().manual_seed(0)
  target = (torch.randn(2, 256, 513, 1, generator=g).abs() * 0.1).cuda()
  x = (target + torch.randn(2, 256, 513, 1, generator=g).cuda() * 0.02)
  full = TrainEngine(spec, 32, {k: v.clone() for k, v in P.items()}, 2, math=N.MATH_FP32)
  masks = {k: (torch.rand(full.G.dropout_shape(k), generator=g) < 0.5).to(torch.uint8).cuda()
           for k in (5, 4)}
  halves = [TrainEngine(spec, 32, {k: v.clone() for k, v in P.items()}, 1, math=N.MATH_FP32)
            for _ in range(2)]
  for step in ('d_step', 'g_step'):
    getattr(full, step)(x, target, dropout=masks, apply=False)
    acc = None
    for r, eng in enumerate(halves):
      m = {k: v[r:r + 1].contiguous() for k, v in masks.items()}
      getattr(eng, step)(x[r:r + 1].contiguous(), target[r:r + 1].contiguous(), dropout=m, apply=False)
      acc = eng.flat.g.clone() if acc is None else acc + eng.flat.g
    lo, hi = full.flat.dis_range() if step == 'd_step' else full.flat.gen_range()
    assert _rel(acc[lo:hi] * 0.5, full.flat.g[lo:hi]) < 1e-4


def test_streaming_inference_matches_serial_calls():
  """MelToMag.run_stream (copies overlapped with the neighbouring batches, two graph / buffer sets)
  returns exactly what back-to-back __call__s return."""
  import torch
  from advoc_b200.infer import MelToMag
  from advoc_b200.model import AdvocSmall, Modes
  m = AdvocSmall(Modes.INFER)
  m.init_params(seed=0)
  eng = MelToMag(m, 2, 'linear', dropout=None, use_graph=True)
  g = torch.Generator().manual_seed(9)
  batches = [torch.randn(2, 256, 80, generator=g).abs() for _ in range(5)]
  want = [eng(b).clone() for b in batches]
  got = [o.clone() for o in eng.run_stream(batches)]
  assert len(got) == 5
  for a, b in zip(got, want):
    assert torch.equal(a, b)
  assert list(eng.run_stream([])) == []


@pytest.mark.parametrize('math', ['f16', 'auto'])
def test_graph_replay_resamples_dropout(math):
  """The captured inference graph reads its dropout seed from a device counter it bumps itself, so
  every replay draws fresh masks (the reference resamples on every sess.run, advoc_model.py:144-149)
  and call k of a graphed engine equals call k of an eager engine, seed for seed."""
  from advoc_b200 import _native as N
  from advoc_b200 import infer
  from advoc_b200.model import AdvocSmall, Modes
  from oracle import nets_torch as O
  P = O.init_params(O.SMALL, seed=0)
  mm = N.MATH_F16 if math == 'f16' else N.MATH_AUTO
  model = AdvocSmall(Modes.INFER, params={k: v.cuda() for k, v in P.items()})
  g = torch.Generator().manual_seed(3)
  mel = torch.randn(2, 256, 80, generator=g).abs()
  eng_g = infer.MelToMag(model, 2, 'linear', dropout='rng', use_graph=True, math=mm)
  eng_e = infer.MelToMag(model, 2, 'linear', dropout='rng', use_graph=False, math=mm)
  outs_g = [eng_g(mel).clone() for _ in range(3)]
  outs_e = [eng_e(mel).clone() for _ in range(3)]
  assert not torch.equal(outs_g[0], outs_g[1]) and not torch.equal(outs_g[1], outs_g[2])
  for a, b in zip(outs_g, outs_e):
    assert torch.equal(a, b)
  # the streaming API (two captured graphs sharing the counter) resamples too
  eng_s = infer.MelToMag(model, 2, 'linear', dropout='rng', use_graph=True, math=mm)
  outs_s = [o.clone() for o in eng_s.run_stream([mel] * 4)]
  assert not torch.equal(outs_s[0], outs_s[2]) and not torch.equal(outs_s[1], outs_s[3])
  for a, b in zip(outs_s[:3], outs_e):
    assert torch.equal(a, b)


def test_loader_melspec_extract_and_paired_audio_slices(tmp_path):
  """advoc/loader.py:103-115 (extract_type='melspec': what models/melspecgan/train.py:20-42 trains on)
  and :133-186 (feature and audio slices cut in parallel, the audio slice slice_len * nhop samples long)
  through the reference's own entry point `decode_extract_and_batch`."""
  from advoc_b200 import audioio, loader
  from oracle import spectral_np as O
  rng = np.random.RandomState(0)
  wav = rng.uniform(-0.5, 0.5, (40000, 1, 1)).astype(np.float32)
  f, a = loader.extract_slices(wav, 'melspec', 64, slice_pad_end=True)
  mel = O.waveform_to_r9y9_melspec(wav).astype(np.float32)           # [157, 80, 1]
  n = -(-mel.shape[0] // 64)
  assert f.shape == (n, 64, 80, 1) and a.shape == (n, 64 * 256, 1, 1)
  ref = np.zeros((n * 64, 80, 1), np.float32)
  ref[:mel.shape[0]] = mel
  assert _rel(f.reshape(-1, 80, 1), ref) < 1e-4
  flat = np.zeros(n * 64 * 256, np.float32)
  flat[:40000] = wav[:, 0, 0]
  assert np.array_equal(a.cpu().numpy().reshape(-1), flat)
  # no padding: whole slices only, both streams the same count
  f2, a2 = loader.extract_slices(wav, 'magspec', 64, slice_pad_end=False)
  assert f2.shape == (2, 64, 513, 1) and a2.shape == (2, 64 * 256, 1, 1)
  # random offset: the audio offset follows the feature offset (loader.py:160-165)
  class _R(object):
    def randint(self, n):
      return 5
  f3, a3 = loader.extract_slices(wav, 'magspec', 64, slice_pad_end=False, slice_randomize_offset=True, rng=_R())
  assert np.array_equal(a3[0, :, 0, 0].cpu().numpy(), wav[5 * 256:5 * 256 + 64 * 256, 0, 0])
  assert _rel(f3[0], f2.reshape(-1, 513, 1)[5:69]) < 1e-5
  # the reference's entry point over files on disk
  fps = []
  for i in range(3):
    fp = str(tmp_path / ('f%d.wav' % i))
    audioio.save_as_wav(fp, 22050, rng.uniform(-0.5, 0.5, (30000 + 4000 * i, 1, 1)).astype(np.float32))
    fps.append(fp)
  got = list(loader.decode_extract_and_batch(fps, batch_size=2, slice_len=64, decode_fastwav=True,
                                             extract_type='melspec', slice_pad_end=True))
  n_slices = sum(-(-(-(-(30000 + 4000 * i) // 256)) // 64) for i in range(3))
  assert len(got) == n_slices // 2
  for xf, xa in got:
    assert xf.shape == (2, 64, 80, 1) and xa.shape == (2, 64 * 256, 1, 1) and xf.is_cuda
  with pytest.raises(ValueError):
    list(loader.decode_extract_and_batch(fps, 2, 64, extract_type='bogus'))
