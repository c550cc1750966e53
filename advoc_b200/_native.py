"""ctypes binding of libadvoc_b200.so (the C-ABI in include/advoc_b200.h).

This is the only place Python touches the native library.  There is no CPU fallback:
if the shared object is missing `lib()` raises, and every wrapper raises on a non-zero
status with the reference's exception vocabulary (BAD_* -> ValueError, UNSUPPORTED ->
NotImplementedError, CUDA -> RuntimeError).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libadvoc_b200.so')

OK, BAD_ARG, BAD_SHAPE, BAD_ALIGN, UNSUPPORTED, WORKSPACE_TOO_SMALL, CUDA_ERROR = range(7)
ACT_NONE, ACT_LRELU, ACT_RELU, ACT_SIGMOID, ACT_TANH = range(5)
MATH_AUTO, MATH_FP32, MATH_TF32, MATH_F16 = range(4)
DT_F32, DT_F16 = range(2)


class ConvDesc(C.Structure):
  """advoc_conv_desc"""
  _fields_ = [(n, C.c_int) for n in
              ('N', 'H', 'W', 'Cin', 'Cout', 'kh', 'kw', 'sh', 'sw', 'pad_t', 'pad_l', 'Ho', 'Wo',
               'math')]


class Epilogue(C.Structure):
  """advoc_epilogue"""
  _fields_ = [
      ('d_bias', C.c_void_p),
      ('act0', C.c_int), ('act1', C.c_int),
      ('alpha', C.c_float),
      ('d_out0', C.c_void_p), ('ld0', C.c_int), ('c_off0', C.c_int),
      ('d_out1', C.c_void_p), ('ld1', C.c_int), ('c_off1', C.c_int),
      ('store_w', C.c_int),
      ('d_dropout_mask', C.c_void_p),
      ('keep_prob', C.c_float),
      ('seed', C.c_uint64),
      ('round_tf32', C.c_int),
      ('accumulate', C.c_int),
      ('d_gate', C.c_void_p), ('ld_gate', C.c_int), ('c_off_gate', C.c_int),
      ('gate_act', C.c_int), ('gate_split', C.c_int),
      ('gate_scale0', C.c_float), ('gate_scale1', C.c_float),
      ('d_seed', C.c_void_p),
      ('out0_dtype', C.c_int), ('out1_dtype', C.c_int),
      ('out0_row_pad', C.c_int), ('reserved0', C.c_int),
  ]


_P = C.c_void_p
_I = C.c_int
_L = C.c_long
_F = C.c_float

# name -> (restype, argtypes); must list every symbol include/advoc_b200.h declares
SIGNATURES = {
    'advoc_version': (_I, []),
    'advoc_last_error': (_I, [C.c_char_p, C.c_size_t]),
    'advoc_launch_count': (C.c_ulonglong, []),
    'advoc_device_arch': (_I, [C.POINTER(_I)]),
    'advoc_num_frames': (_I, [_I, _I, _I, _I]),
    'advoc_stft_f32': (_I, [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    'advoc_melspec_f32': (_I, [_P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _I, _F, _F, _P, _P]),
    'advoc_mel_ranges': (_I, [_P, _I, _I, _P, _P]),
    'advoc_matmul_lastdim_f32': (_I, [_P, _P, _P, _L, _I, _I, _I, _F, _F, _P]),
    'advoc_conv2d_fwd': (_I, [C.POINTER(ConvDesc), _P, _I, _P, C.POINTER(Epilogue), _P]),
    'advoc_pack_filter': (_I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    'advoc_debug_flags': (_I, [C.POINTER(C.c_uint)]),
    'advoc_debug_peek': (_I, [C.POINTER(C.c_uint)]),
    'advoc_istft_frames_f32': (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _P]),
    'advoc_overlap_add_f32': (_I, [_P, _I, _I, _I, _I, _P, _P]),
    'advoc_griffin_lim_iter_f32': (_I, [_P, _L, _P, _I, _I, _I, _I, _P, _P, _P, _P]),
    'advoc_conv2d_wgrad': (_I, [C.POINTER(ConvDesc), _P, _I, _P, _I, _P, _P]),
    'advoc_bias_grad': (_I, [_P, _I, _L, _I, _P, _P]),
    'advoc_gan_logloss': (_I, [_P, _P, _L, _I, _F, _P, _P, _P, _P]),
    'advoc_l1_loss': (_I, [_P, _I, _I, _P, _L, _F, _P, _P, _I, _P]),
    'advoc_adam_tf_step': (_I, [_P, _P, _P, _P, _L, _F, _F, _F, _F, _L, _F, _P]),
    'advoc_adam_tf_step_dev': (_I, [_P, _P, _P, _P, _L, _P, _F, _F, _F, _F, _P]),
    'advoc_conv2d_path': (_I, [C.POINTER(ConvDesc), _I, _I]),
    'advoc_conv2d_kernel': (_I, [C.POINTER(ConvDesc), _I, _I, _I]),
    'advoc_conv2d_tile_n': (_I, [C.POINTER(ConvDesc), _I, _I, _I]),
    'advoc_gemm_f32': (_I, [_P, _I, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    'advoc_bn_stats': (_I, [_P, _I, _L, _I, _P, _P]),
    'advoc_bn_apply': (_I, [_P, _I, _L, _I, _P, _P, _P, _F, _I, _F, _P, _I, _I, _P]),
    'advoc_bn_inference': (_I, [_P, _I, _L, _I, _P, _P, _P, _P, _F, _I, _F, _P, _I, _I, _P]),
    'advoc_bn_moving_update': (_I, [_P, _L, _I, _F, _P, _P, _P]),
    'advoc_bn_backward': (_I, [_P, _I, _P, _I, _P, _I, _L, _I, _P, _P, _F, _I, _F, _P, _P, _I, _I, _P]),
    'advoc_tanh_backward': (_I, [_P, _P, _P, _L, _P]),
    'advoc_gan_logit_loss': (_I, [_P, _P, _I, _I, _P, _P, _P, _P]),
    'advoc_gp_seed': (_I, [_P, _I, _L, _F, _P, _P, _I, _P]),
    'advoc_bn_gp': (_I, [_P, _P, _P, _P, _L, _I, _P, _P, _F, _F, _P, _P, _P, _I, _P]),
    'advoc_conv2d_transpose_fwd': (_I, [C.POINTER(ConvDesc), _P, _I, _P, C.POINTER(Epilogue), _P]),
}

_lib = None


def lib():
  """The loaded library; raises (never falls back) if it has not been built."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError(
          'advoc_b200: %s is missing -- build it with `make -C advoc_b200/csrc` or '
          '`python -c "import __graft_entry__ as g; g.build()"`; there is no CPU fallback'
          % LIB_PATH)
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
      fn = getattr(l, name)
      fn.restype = res
      fn.argtypes = args
    _lib = l
  return _lib


def last_error():
  buf = C.create_string_buffer(512)
  lib().advoc_last_error(buf, 512)
  return buf.value.decode('utf-8', 'replace')


def check(status):
  if status == OK:
    return
  msg = last_error()
  if status in (BAD_ARG, BAD_SHAPE, BAD_ALIGN, WORKSPACE_TOO_SMALL):
    raise ValueError(msg)
  if status == UNSUPPORTED:
    raise NotImplementedError(msg)
  raise RuntimeError(msg)


def call(name, *args):
  check(getattr(lib(), name)(*args))


def debug_flags():
  v = C.c_uint(0)
  call('advoc_debug_flags', C.byref(v))
  return v.value


def raise_if_aborted(where=''):
  """Cheap check (pinned host word, no CUDA call) for the host engines' sync points: raises if a
  tcgen05 pipeline wait timed out since the flag was last cleared -- every later tensor-core
  launch then returned without writing its outputs."""
  v = C.c_uint(0)
  call('advoc_debug_peek', C.byref(v))
  if v.value:
    raise RuntimeError('advoc_b200: a tcgen05 pipeline wait timed out (code %d)%s; results since then are '
                       'invalid -- clear with _native.debug_flags()' % (v.value, ' in ' + where if where else ''))


def launch_count():
  return int(lib().advoc_launch_count())


def device_arch():
  a = _I(0)
  call('advoc_device_arch', C.byref(a))
  return a.value
