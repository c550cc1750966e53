"""The C-ABI boundary without a GPU: the library loads, exports every symbol the header
declares, the ctypes table covers them all, and host-side argument validation maps onto the
reference's exception vocabulary.  No kernel is launched here."""
import ctypes as C
import os
import re

import pytest

from advoc_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  src = open(os.path.join(ROOT, 'include', 'advoc_b200.h')).read()
  return sorted(set(re.findall(r'ADVOC_API\s+[\w\s\*]+?\b(advoc_\w+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
  names = _declared()
  assert len(names) >= 10
  lib = C.CDLL(N.LIB_PATH)
  for n in names:
    assert hasattr(lib, n), 'missing export ' + n
  assert sorted(N.SIGNATURES) == names


def test_version_and_error_buffer():
  assert N.lib().advoc_version() == 100
  buf = C.create_string_buffer(64)
  assert N.lib().advoc_last_error(buf, 64) == N.OK
  assert N.lib().advoc_last_error(None, 0) == N.BAD_ARG


def test_num_frames_rule():
  f = N.lib().advoc_num_frames
  assert f(16000, 1024, 256, 1) == 63     # reference tests/test_spectral.py:31-33
  assert f(16000, 1024, 256, 0) == 60     # :35-36
  assert f(16384, 1024, 256, 1) == 64
  assert f(22050, 1024, 256, 1) == 87
  assert f(0, 1024, 256, 1) == 0
  assert f(100, 1024, 256, 0) == 1
  assert f(16000, 1024, 256, 2) == 59 and f(100, 1024, 256, 2) == 0    # tf.contrib.signal.stft(pad_end=False)


def test_bad_arguments_raise_without_touching_the_gpu():
  with pytest.raises(ValueError):
    N.call('advoc_stft_f32', None, 1, 100, 1, 1024, 256, 1, None, None, None, None, None)
  with pytest.raises(ValueError):
    N.call('advoc_matmul_lastdim_f32', None, None, None, 4, 0, 3, 0, -100.0, 20.0, None)
  d = N.ConvDesc(1, 8, 8, 4, 4, 4, 4, 2, 2, 1, 1, 4, 4, N.MATH_FP32)
  with pytest.raises(ValueError):
    N.call('advoc_conv2d_fwd', C.byref(d), None, 4, None, None, None)
  bad = N.ConvDesc(1, 8, 8, 0, 4, 4, 4, 2, 2, 1, 1, 4, 4, N.MATH_FP32)
  with pytest.raises(ValueError):
    N.call('advoc_conv2d_fwd', C.byref(bad), None, 4, None, None, None)
  assert 'shape' in N.last_error()


def test_struct_layouts_match_the_header():
  # advoc_conv_desc: 14 ints; advoc_epilogue: see include/advoc_b200.h
  assert C.sizeof(N.ConvDesc) == 14 * 4
  assert N.Epilogue.d_out0.offset == 24 and N.Epilogue.d_out1.offset == 40
  assert N.Epilogue.d_dropout_mask.offset == 64 and N.Epilogue.seed.offset == 80
  assert N.Epilogue.d_gate.offset == 96 and N.Epilogue.gate_scale1.offset == 124
  assert N.Epilogue.d_seed.offset == 128 and N.Epilogue.out0_dtype.offset == 136
  assert N.Epilogue.out0_row_pad.offset == 144
  assert C.sizeof(N.Epilogue) == 152


def _prototypes():
  """{name: [C parameter type strings]} parsed from the header (comments stripped)."""
  src = open(os.path.join(ROOT, 'include', 'advoc_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', ' ', src, flags=re.S)
  out = {}
  for name, args in re.findall(r'ADVOC_API\s+[\w\s\*]+?\b(advoc_\w+)\s*\(([^;]*?)\)\s*;', src, flags=re.S):
    args = ' '.join(args.split())
    params = [] if args in ('', 'void') else [a.strip() for a in args.split(',')]
    out[name] = [re.sub(r'\s*\b\w+$', '', p).strip() if not p.endswith('*') else p for p in params]
  return out


def test_ctypes_signatures_match_the_header_prototypes():
  """Every entry of _native.SIGNATURES has the header's parameter count, and each ctypes type is of the
  header parameter's kind (pointer / long / int / float / size_t): a silent mismatch here would shift
  every later argument of a call."""
  protos = _prototypes()
  assert sorted(protos) == sorted(N.SIGNATURES)
  def kind_of_ctype(ct):
    if ct in (C.c_void_p, C.c_char_p) or (isinstance(ct, type) and issubclass(ct, C._Pointer)):
      return 'ptr'
    if ct in (C.c_float, C.c_double):
      return 'f%d' % (8 * C.sizeof(ct))
    signed = ct(-1).value < 0
    return '%s%d' % ('i' if signed else 'u', 8 * C.sizeof(ct))

  def kind_of_c(t):
    if '*' in t:
      return 'ptr'
    t = t.replace('const', '').strip()
    return {'int': 'i32', 'long': 'i64', 'long long': 'i64', 'float': 'f32', 'double': 'f64', 'size_t': 'u64',
            'unsigned int': 'u32', 'unsigned': 'u32', 'unsigned long long': 'u64', 'unsigned long': 'u64'}.get(t, t)

  for name, (restype, argtypes) in N.SIGNATURES.items():
    want = protos[name]
    assert len(argtypes) == len(want), (name, len(argtypes), want)
    for i, (ct, c_decl) in enumerate(zip(argtypes, want)):
      assert kind_of_ctype(ct) == kind_of_c(c_decl), (name, i, ct, c_decl)
