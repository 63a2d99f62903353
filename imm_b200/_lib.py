"""ctypes binding of libimm_b200.so (the C ABI declared in include/imm_b200.h).

The product path has NO CPU fallback: if the shared library is missing or a CUDA device is absent,
calls fail loudly.  PyTorch is used only for device memory and streams."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('IMMB_LIB', os.path.join(_HERE, 'lib', 'libimm_b200.so'))     # IMMB_LIB: A/B builds (development)

ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2
PREC_TF32X3, PREC_TF32, PREC_TF32X2, PREC_F16X3, PREC_F16X2 = 0, 1, 2, 3, 4
EPI_BIAS, EPI_BIAS_RELU = 0, 1
XLAYOUT_NHWC, XLAYOUT_ROWWIN4 = 0, 1


class ImmbError(RuntimeError):
  pass


class ConvDesc(ctypes.Structure):
  _fields_ = ([(n, ctypes.c_int32) for n in
               ('N', 'H', 'W', 'Cin', 'Cout', 'kh', 'kw', 'stride', 'Ho', 'Wo', 'pad_t', 'pad_l',
                'x_cstride', 'y_cstride', 'cin_pad', 'epilogue', 'precision', 'engine', 'x_layout', 'reserved_')] +
              [(n, ctypes.c_void_p) for n in ('x_scale', 'y_scale', 'w_scale')])     # H16 scale records (NULL: TF32)


class ReduceItem(ctypes.Structure):
  """immb_reduce_item (include/imm_b200.h)."""
  _fields_ = ([(n, ctypes.c_void_p) for n in ('partials', 'out_f')] +
              [(n, ctypes.c_int32) for n in ('nblocks', 'nvals', 'block0', 'reserved_')])


class PackItem(ctypes.Structure):
  """immb_pack_item (include/imm_b200.h)."""
  _fields_ = ([(n, ctypes.c_void_p) for n in ('w', 'wp_hi', 'wp_lo', 'wh_hi', 'wh_lo', 'amax', 'rec')] +
              [(n, ctypes.c_int32) for n in ('taps', 'Cin', 'Cout', 'cin_pad', 'cout_pad', 'kind', 'block0', 'nblocks')])


_P, _I, _L, _F, _Z = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
_D = ctypes.POINTER(ConvDesc)

# name -> argtypes (everything returns int unless listed in _RESTYPES)
_SIGS = {
  'immb_conv_engine_for': [_D, _I],
  'immb_conv2d_fwd': [_D, _P, _P, _P, _P, _P, _P, _P, _P, _P],
  'immb_conv2d_dgrad': [_D, _P, _P, _P, _P, _P, _P, _P],
  'immb_conv2d_fwd_stats_rows': [_D],
  'immb_conv2d_fwd_bnstats': [_D, _P, _P, _P, _P, _P, _P, _P, _Z, _P],
  'immb_bn_stats_from_partials': [_P, _I, _I, _P, _P],
  'immb_conv2d_dgrad_stats_rows': [_D],
  'immb_conv2d_dgrad_bnreduce': [_D, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _P, _Z, _P],
  'immb_conv2d_dgrad_relu': [_D, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P],
  'immb_conv2d_dgrad_relu_supported': [_D],
  'immb_conv2d_wgrad_workspace': [_D],
  'immb_conv2d_wgrad': [_D, _P, _P, _P, _P, _P, _P, _Z, _P],
  'immb_pack_weights': [_P, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P],
  'immb_split_planes': [_P, _P, _P, _L, _P, _P],
  'immb_bn_scratch_elems': [_L, _I],
  'immb_bn_stats': [_P, _L, _I, _I, _P, _P, _Z, _P],
  'immb_bn_finalize': [_P, _L, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P],
  'immb_bn_apply': [_P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P, _I, _P, _P],
  'immb_upsample2x_bwd': [_P, _I, _I, _I, _I, _I, _P, _P],
  'immb_bn_bwd_reduce': [_P, _I, _P, _I, _L, _I, _P, _P, _P, _P, _I, _P, _P, _Z, _P],
  'immb_bn_bwd_apply': [_P, _I, _P, _I, _L, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P, _P, _Z, _P, _P, _I, _P],
  'immb_bn_bwd_apply_blocks': [_L, _I],
  'immb_reduce_partials_multi': [_P, _I, _I, _P],
  'immb_bn_finalize_partials': [_P, _I, _L, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
  'immb_bias_grad': [_P, _P, _I, _L, _I, _P, _P, _P],
  'immb_cast_d2f': [_P, _P, _L, _P],
  'immb_softargmax_gauss_fwd': [_P, _I, _I, _I, _I, _F, _P, _P, _P, _I, _P, _P, _I, _I, _P, _P],
  'immb_softargmax_gauss_bwd': [_P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _F, _P, _I, _P],
  'immb_gaussian_maps': [_P, _I, _I, _I, _F, _P, _P],
  'immb_vgg_prologue': [_P, _P, _I, _I, _I, _I, _P, _P, _P],
  'immb_vgg_conv1_1_fused': [_P, _P, _I, _I, _I, _P, _P, _I, _P, _P, _I, _P, _P],
  'immb_stage_image_rowwin': [_P, _I, _I, _I, _P, _P, _P],
  'immb_pack_weights_rowwin': [_P, _I, _P, _P, _P],
  'immb_maxpool2x2_fwd': [_P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P],
  'immb_maxpool2x2_fwd_levelsum': [_P, _P, _I, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P],
  'immb_maxpool2x2_bwd_combine': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P],
  'immb_maxpool2x2_bwd': [_P, _P, _P, _I, _I, _I, _I, _P, _P],
  'immb_perceptual_level_sum': [_P, _P, _I, _P, _P, _I, _I, _I, _I, _I, _P, _I, _P, _P, _P],
  'immb_perceptual_finalize': [_P, _P, _I, _P, _I, _P, _P, _P, _P],
  'immb_vgg_bwd_combine': [_P, _P, _P, _P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P],
  'immb_vgg_conv1_1_bwd_fused': [_P, _P, _P, _I, _P, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _P],
  'immb_pred_grad': [_P, _P, _I, _P, _P, _P, _I, _I, _I, _P, _P, _P],
  'immb_resize_ac_fwd': [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P],
  'immb_resize_ac_bwd': [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P],
  'immb_tps_warp': [_P, _I, _I, _I, _I, _P, _I, _I, _P, _P],
  'immb_adam_norms': [_P, _P, _L, _P, _P, _P, _I, _P, _F, _P, _P, _P],
  'immb_adam_apply': [_P, _P, _P, _P, _L, _P, _P, _P, _I, _P, _F, _P, _F, _F, _F, _F, _F, _P, _P],
  'immb_adam_apply_dev': [_P, _P, _P, _P, _L, _P, _P, _P, _I, _P, _F, _P, _F, _P, _F, _F, _F, _P, _P],
  'immb_total_loss': [_P, _P, _P, _I, _P, _P, _P, _P],
  'immb_scale_update': [_P, _I, _P, _P],
  'immb_multi_amax': [_P, _P, _P, _P, _I, _P, _P],
  'immb_pack_weights_multi': [_P, _I, _I, _P],
  'immb_crc32c': [_P, _Z, ctypes.c_uint32],
}
# trailing H16 scale-record / amax arguments (just before `stream`): callers that work on fp32 TF32 planes may omit
# them -- call() fills them with NULL
_OPTIONAL_TAIL = {'immb_conv2d_dgrad_relu': 1, 'immb_pack_weights': 2, 'immb_split_planes': 1, 'immb_bn_apply': 1,
                  'immb_bn_bwd_apply': 3, 'immb_bias_grad': 1, 'immb_softargmax_gauss_fwd': 1, 'immb_vgg_conv1_1_fused': 1,
                  'immb_maxpool2x2_fwd': 2, 'immb_maxpool2x2_fwd_levelsum': 2, 'immb_maxpool2x2_bwd_combine': 2,
                  'immb_perceptual_level_sum': 1, 'immb_vgg_bwd_combine': 2, 'immb_vgg_conv1_1_bwd_fused': 2,
                  'immb_resize_ac_fwd': 2, 'immb_adam_apply': 1, 'immb_adam_apply_dev': 1, 'immb_total_loss': 1}
_RESTYPES = {'immb_bn_bwd_apply_blocks': ctypes.c_int, 'immb_conv2d_wgrad_workspace': _Z, 'immb_bn_scratch_elems': _Z, 'immb_crc32c': ctypes.c_uint32}

_lib = None


def lib():
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise ImmbError('libimm_b200.so not found at %s -- run `python -c "import __graft_entry__ as g; g.build()"` '
                      '(there is no CPU fallback)' % LIB_PATH)
    l = ctypes.CDLL(LIB_PATH)
    l.immb_version.restype = ctypes.c_int
    l.immb_last_error.restype = ctypes.c_char_p
    l.immb_launch_count.restype = ctypes.c_int64
    for name, argtypes in _SIGS.items():
      fn = getattr(l, name)
      fn.argtypes = argtypes
      fn.restype = _RESTYPES.get(name, ctypes.c_int)
    _lib = l
  return _lib


def exported_symbols():
  return ['immb_version', 'immb_last_error', 'immb_launch_count'] + sorted(_SIGS.keys())


def ptr(t):
  """device pointer of a tensor (or None).  Tensors must be fp32/fp64/int CUDA tensors laid out as the
  kernel expects; only data_ptr() crosses the ABI."""
  if t is None:
    return None
  if isinstance(t, int):
    return t
  return t.data_ptr()


def stream_ptr():
  return torch.cuda.current_stream().cuda_stream


PROFILE = None          # development aid: when a list, every call is timed with CUDA events -> (name, tag, ms)
TAG = ''


def call(name, *args):
  """Calls a C-ABI function; tensors are converted to raw pointers; raises ImmbError on failure."""
  if PROFILE is not None and name not in _RESTYPES:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = _call(name, *args)
    e1.record()
    PROFILE.append((name, TAG, e0, e1, conv_info(name, args[0]) if args and isinstance(args[0], ConvDesc) else None))
    return rc
  return _call(name, *args)


def conv_info(name, d):
  """Algorithmic work of one conv call (2*MACs) and which kernel family serves it (development /
  bench.py roofline attribution; mirrors conv_tc2_eligible / conv_tc2_rowwin_eligible in csrc/conv_tc2.cu)."""
  flops = 2.0 * d.N * d.Ho * d.Wo * d.Cout * d.kh * d.kw * d.Cin
  halo = d.kh == 3 and d.stride == 1 and d.H % 16 == 0 and d.W % 16 == 0 and d.x_layout == XLAYOUT_NHWC
  first = d.x_layout == XLAYOUT_ROWWIN4 and d.H % 16 == 0 and d.W % 8 == 0
  wgrad = name.endswith('wgrad')
  if wgrad:
    kernel = 'conv_tc2_wgrad_kernel' if (d.kh == 3 and d.stride == 1 and d.H % 4 == 0 and d.W % 8 == 0
                                         and d.x_layout == XLAYOUT_NHWC) else 'conv_tc_wgrad_kernel'
  else:
    kernel = 'conv_tc2_pair_kernel' if (halo or (first and 'fwd' in name)) else 'conv_tc_kernel'
  passes = {PREC_TF32X3: 3, PREC_TF32: 1, PREC_TF32X2: 2, PREC_F16X3: 3, PREC_F16X2: 2}[d.precision]
  f16 = d.precision in (PREC_F16X3, PREC_F16X2)
  if not (kernel == 'conv_tc2_pair_kernel') and passes == 2:
    passes = 3                      # only the pair kernel has a 2-pass variant
  if f16 and wgrad:
    kernel = 'conv_tc2_wgrad16_kernel'
  return {'flops': flops, 'kernel': kernel, 'passes': passes, 'kind': 'f16' if f16 else 'tf32'}


def _call(name, *args):
  l = lib()
  conv = []
  for a in args:
    if isinstance(a, torch.Tensor):
      if not a.is_cuda:
        raise ImmbError('%s: got a CPU tensor; the CUDA path has no CPU fallback' % name)
      conv.append(a.data_ptr())
    elif isinstance(a, ConvDesc):
      conv.append(ctypes.byref(a))
    else:
      conv.append(a)
  k = _OPTIONAL_TAIL.get(name, 0)
  missing = len(_SIGS[name]) - len(conv) if name in _SIGS else 0
  if k and 0 < missing <= k:
    sig = _SIGS[name]
    fill = [None if sig[len(conv) - 1 + j] is _P else 0 for j in range(missing)]
    conv = conv[:-1] + fill + conv[-1:]
  rc = getattr(l, name)(*conv)
  if name in _RESTYPES:
    return rc
  if rc != 0:
    raise ImmbError('%s failed (%d): %s' % (name, rc, l.immb_last_error().decode()))
  return rc


def launch_count():
  return int(lib().immb_launch_count())
