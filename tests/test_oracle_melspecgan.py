"""CPU checks of the MelspecGAN oracle (oracle/melspecgan_torch.py): layer shapes and parameter
counts of SURVEY.md appendix A.3, the TF SAME-padding identities of appendix B for k5 s2, batch
norm against torch's own, and that both reference losses (train.py:76-111) differentiate."""
import torch
import torch.nn.functional as F

from oracle import melspecgan_torch as M


def test_shapes_and_param_counts():
  P = M.init_params(seed=0)
  g = torch.Generator().manual_seed(1)
  z = torch.randn(3, M.Z_DIM, generator=g)
  G_z, layers = M.generator(P, z, return_layers=True)
  assert [tuple(l.shape[1:]) for l in layers] == [(4, 5, 512), (8, 10, 256), (16, 20, 128), (32, 40, 64),
                                                  (64, 80, 1)]
  assert float(G_z.abs().max()) < 1.0
  out, dl = M.discriminator(P, G_z, return_layers=True)
  assert out.shape == (3,)
  assert [tuple(l.shape[1:]) for l in dl] == [(32, 40, 64), (16, 20, 128), (8, 10, 256), (4, 5, 512)]
  n_g = sum(P[n].numel() for n in M.g_names(P) if n.endswith('/W') or n.endswith('/b'))
  n_d = sum(P[n].numel() for n in M.d_names(P) if n.endswith('/W') or n.endswith('/b'))
  assert n_g == 5337089 and n_d == 4313601


def test_k5_same_padding_identities():
  """conv SAME k5 s2 pads (1, 2); conv_transpose SAME is its adjoint (== the input gradient)."""
  g = torch.Generator().manual_seed(2)
  x = torch.randn(2, 8, 10, 3, generator=g, requires_grad=True)
  W = torch.randn(5, 5, 3, 4, generator=g)
  y = M.conv5(x, W, None)
  assert y.shape == (2, 4, 5, 4)
  dy = torch.randn(y.shape, generator=g)
  (gx,) = torch.autograd.grad(y, x, dy)
  # deconv5 takes the conv_transpose layout [kh,kw,out,in] = the conv's HWIO with roles swapped
  assert torch.allclose(M.deconv5(dy, W, None), gx, atol=1e-5)


def test_batchnorm_matches_torch():
  g = torch.Generator().manual_seed(3)
  x = torch.randn(4, 6, 5, 8, generator=g) * 3 + 1
  gamma, beta = torch.randn(8, generator=g), torch.randn(8, generator=g)
  ref = F.batch_norm(x.permute(0, 3, 1, 2), None, None, gamma, beta, training=True, eps=M.BN_EPS).permute(0, 2, 3, 1)
  assert torch.allclose(M.batchnorm(x, gamma, beta), ref, atol=1e-5)


def test_losses_differentiate():
  P = {n: t.clone().requires_grad_(True) for n, t in M.init_params(seed=0, dim=8).items()}
  g = torch.Generator().manual_seed(4)
  z = torch.randn(4, M.Z_DIM, generator=g)
  x = torch.rand(4, 64, 80, 1, generator=g) * 2 - 1
  for kind in ('dcgan', 'wgangp'):
    l = M.losses_dim(P, z, x, 8, kind, alpha=torch.rand(4, 1, 1, 1, generator=g))
    gd = torch.autograd.grad(l['D_loss'], [P[n] for n in M.d_names(P)], retain_graph=True, allow_unused=True)
    gg = torch.autograd.grad(l['G_loss'], [P[n] for n in M.g_names(P)], allow_unused=True)
    assert all(t is not None and torch.isfinite(t).all() for t in gd)
    assert all(t is not None and torch.isfinite(t).all() for t in gg)


def test_moving_average_update_rule():
  """One training-mode evaluation from the TF initial values (0, 1): moving -= (moving - batch) * 0.01 with
  the Bessel-corrected batch variance; training=False then reads them and a larger batch does not matter."""
  import torch
  from oracle import melspecgan_torch as M
  P = M.init_params(seed=1, dim=8)
  g = torch.Generator().manual_seed(2)
  z = torch.randn(5, M.Z_DIM, generator=g)
  moving = M.init_moving(8)
  M.generator(P, z, 8, moving=moving)
  x0 = (z @ P['G/z_proj/W'] + P['G/z_proj/b']).reshape(-1, 4, 5, 64).reshape(-1, 64)
  assert torch.allclose(moving['G/batch_normalization/moving_mean'], 0.01 * x0.mean(0), atol=1e-7)
  assert torch.allclose(moving['G/batch_normalization/moving_variance'], 0.99 + 0.01 * x0.var(0, unbiased=True),
                        atol=1e-7)
  one = M.generator(P, z[:1], 8, moving=moving, training=False)
  many = M.generator(P, z, 8, moving=moving, training=False)
  assert torch.allclose(one, many[:1], atol=1e-6)      # no batch coupling in the inference graph


def _graph_fixture():
  import json
  import os
  fp = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'melspecgan_infer_graph.json')
  with open(fp) as f:
    return json.load(f)


def test_generator_matches_the_reference_metagraph():
  """models/melspecgan/infer.meta (the reference's exported inference graph; facts extracted by
  scripts/make_melspecgan_graph_fixture.py) pins the generator's variable names / layouts, op order,
  conv_transpose strides + padding + output sizes, the fused batch norm's epsilon and decay, the latent
  size and the feats_denorm tail.  The oracle and the product's host constants must agree with it."""
  import numpy as np
  import torch
  from oracle import melspecgan_torch as M
  fx = _graph_fixture()
  # --- variables: names, shapes, dtype (DT_FLOAT = 1); global_step is the only non-G variable
  ref_vars = {v['name']: tuple(v['shape']) for v in fx['variables']}
  assert all(v['dtype'] == 1 for v in fx['variables'] if v['name'] != 'global_step')
  assert ref_vars.pop('global_step') == ()
  P = M.init_params(seed=0, dim=64)
  mine = {k: tuple(v.shape) for k, v in P.items() if k.startswith('G/')}
  mine.update({k: tuple(v.shape) for k, v in M.init_moving(64).items()})
  assert mine == ref_vars
  # --- op chain from the latent to G_z
  ops = {o['name']: o for o in fx['ops']}
  assert ops['G/z_proj/MatMul']['inputs'] == ['z', 'G/z_proj/W/read']
  assert ops['G/z_proj/MatMul']['attr'] == {'transpose_a': False, 'transpose_b': False}
  assert ops['G/Reshape']['inputs'][0] == 'G/z_proj/BiasAdd' and fx['consts']['G/Reshape/shape'] == [-1, 4, 5, 512]
  prev, sizes = 'G/Reshape', []
  for i in range(4):
    bn = 'G/batch_normalization%s/FusedBatchNorm' % ('' if i == 0 else '_%d' % i)
    stem = bn.rsplit('/', 1)[0]
    assert stem == M.G_BN[i]
    assert ops[bn]['inputs'] == [prev] + [stem + s for s in ('/gamma/read', '/beta/read', '/moving_mean/read',
                                                             '/moving_variance/read')]
    assert ops[bn]['attr']['is_training'] is False and ops[bn]['attr']['data_format'] == 'NHWC'
    assert np.float32(ops[bn]['attr']['epsilon']) == np.float32(M.BN_EPS)
    assert np.float32(fx['consts'][stem + '/Const'][0]) == np.float32(M.BN_MOMENTUM)
    relu = 'G/Relu' + ('' if i == 0 else '_%d' % i)
    assert ops[relu]['inputs'] == [bn]
    up = 'G/upconv_%d' % (i + 1)
    ct = ops[up + '/conv2d_transpose']
    assert ct['op'] == 'Conv2DBackpropInput' and ct['inputs'][1:] == [up + '/W/read', relu]
    assert ct['attr']['strides'] == [1, 2, 2, 1] and ct['attr']['padding'] == 'SAME'
    assert ct['attr']['data_format'] == 'NHWC' and ct['attr']['dilations'] == [1, 1, 1, 1]
    assert ops[up + '/BiasAdd']['inputs'] == [up + '/conv2d_transpose', up + '/b/read']
    sizes.append(tuple(fx['consts'][up + '/conv2d_transpose/output_shape/%d' % k][0] for k in (1, 2, 3)))
    prev = up + '/BiasAdd'
  assert ops['G/Tanh']['inputs'] == ['G/upconv_4/BiasAdd']
  tail = {t['name']: t for t in fx['tail']}
  assert tail['add']['inputs'] == ['G/Tanh', 'add/y'] and tail['mul']['inputs'] == ['add', 'mul/y']
  assert tail['G_z']['inputs'] == ['mul'] and fx['consts']['add/y'] == [1.0] and fx['consts']['mul/y'] == [0.5]
  assert fx['consts']['samp_z/shape/1'] == [M.Z_DIM]
  # --- the oracle's activations have the sizes the graph requests from conv2d_transpose
  with torch.no_grad():
    out, layers = M.generator(P, torch.zeros(2, M.Z_DIM), 64, return_layers=True, moving=M.init_moving(64),
                              training=False)
  assert [tuple(l.shape[1:]) for l in layers[1:]] == sizes and sizes[-1] == (64, 80, 1)
  # --- the product's host-side constants (no device work at import)
  from advoc_b200 import melspecgan as MG
  assert MG.Z_DIM == M.Z_DIM and np.float32(MG.BN_EPS) == np.float32(M.BN_EPS) and list(MG.G_BN) == list(M.G_BN)


def test_graph_fixture_is_reproducible_from_the_reference():
  """In the authoring container (reference checkout present) the committed fixture must be exactly what
  scripts/make_melspecgan_graph_fixture.py extracts; on the GPU box there is no reference: skipped."""
  import importlib.util
  import json
  import os
  import pytest
  src = '/root/reference/models/melspecgan/infer.meta'
  if not os.path.isfile(src):
    pytest.skip('reference checkout not present')
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  spec = importlib.util.spec_from_file_location('mkfix', os.path.join(root, 'scripts', 'make_melspecgan_graph_fixture.py'))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  import tempfile
  with tempfile.TemporaryDirectory() as d:
    mod.DST = os.path.join(d, 'out.json')
    mod.main()
    with open(mod.DST) as f:
      fresh = json.load(f)
  assert fresh == _graph_fixture()
