"""Device-side execution engine of the IMM training step.

Owns every device buffer (PyTorch tensors = memory only) and drives the hand-written sm_100a kernels of
libimm_b200.so through the C ABI (imm_b200/_lib.py).  No torch op computes anything on the hot path and
there is no CPU fallback.  The arithmetic follows the reference call sites cited in include/imm_b200.h and
in the docstrings below (file:line under /root/reference).

Data layout in HBM (per GPU, batch B, NHWC fp32):
  * trainable parameters / gradients / Adam m,v : four flat fp32 buffers, tensors back to back in the
    reference's variable order (TF names), one NCCL all-reduce over the flat gradient buffer;
  * conv operands are "split planes" (hi, lo) = error-compensated TF32 pairs written by the producing kernel;
  * raw conv outputs y (pre-BN) are kept for the BN backward; gradients wrt activations are single fp32.
"""
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib
from ._lib import ConvDesc, call

WD = 1e-5                 # base_model.py:65
INIT_STD = 0.01           # base_model.py:66
PERCEPTUAL_WS = [100.0, 1.6, 2.3, 1.8, 2.8, 100.0]    # imm_model.py:131
VGG_ORDER = [('conv1_1', 64), ('conv1_2', 64), 'pool1', ('conv2_1', 128), ('conv2_2', 128), 'pool2',
             ('conv3_1', 256), ('conv3_2', 256), ('conv3_3', 256), 'pool3',
             ('conv4_1', 512), ('conv4_2', 512), ('conv4_3', 512), 'pool4',
             ('conv5_1', 512), ('conv5_2', 512), ('conv5_3', 512), 'pool5']     # vgg16.py:338-373
ADAM_CHUNK = 2048


def same_pad(in_size, k, stride):
  """TF SAME: out = ceil(in/stride); total = max((out-1)*stride + k - in, 0); before = total//2."""
  out = -(-in_size // stride)
  total = max((out - 1) * stride + k - in_size, 0)
  return total // 2, total - total // 2


def round_up(x, m):
  return (x + m - 1) // m * m


def encoder_spec(n_filters):
  """imm_model.py:182-217 -> [(name, k, stride, cout)]."""
  f = n_filters
  return [('conv_1', 7, 1, f), ('conv_2', 3, 1, f), ('conv_3', 3, 2, 2 * f), ('conv_4', 3, 1, 2 * f),
          ('conv_5', 3, 2, 4 * f), ('conv_6', 3, 1, 4 * f), ('conv_7', 3, 2, 8 * f), ('conv_8', 3, 1, 8 * f)]


def renderer_spec(n_filters_render, final_res, n_final_out, start_res=16):
  """imm_model.py:154-179 -> [(name, cout, batch_norm, relu, upsample_after)]."""
  filters = n_filters_render * 8
  size, conv_id, spec = start_res, 1, []
  while size <= final_res:
    spec.append(('conv_%d' % conv_id, filters, True, True, False))
    if size == final_res:
      spec.append(('conv_%d' % (conv_id + 1), n_final_out, False, False, False))
      break
    spec.append(('conv_%d' % (conv_id + 1), filters, True, True, True))
    size *= 2
    conv_id += 2
    if filters >= 8:
      filters //= 2
  return spec



def _dbg(msg):
  """IMMB_BENCH_TRACE=1: progress lines on stderr (where a multi-rank run stops, if it ever does)."""
  if os.environ.get('IMMB_BENCH_TRACE'):
    import sys
    sys.stderr.write('[engine rank %s] %s\n' % (os.environ.get('RANK', '0'), msg))
    sys.stderr.flush()


class Planes(object):
  """A split tensor: hi (+ optional lo) planes of identical shape [N,H,W,Cs].
  fp32 planes = TF32 pair; fp16 planes ("H16", include/imm_b200.h) carry `scale`, a view of the tensor's scale record
  (int32 {e, amax bits}): value = (hi + lo * 2^-11) * 2^-e."""

  def __init__(self, hi, lo=None, scale=None):
    self.hi, self.lo, self.scale = hi, lo, scale

  @property
  def h16(self):
    return self.scale is not None

  @staticmethod
  def alloc(shape, device, lo=True, zero=False, scale=None):
    f = torch.zeros if zero else torch.empty
    dt = torch.float16 if scale is not None else torch.float32
    return Planes(f(shape, dtype=dt, device=device), f(shape, dtype=dt, device=device) if lo else None, scale)

  def value(self):
    """The tensor the planes represent, as fp32 (host-side convenience: eval outputs, sub-builders, tests)."""
    if self.scale is not None:
      e = int(self.scale[0].item())
      return (self.hi.float() + self.lo.float() * (2.0 ** -11)) * (2.0 ** -e)
    return self.hi if self.lo is None else self.hi + self.lo

  def half(self, i, B):
    """i-th batch half of a [2B,...] tensor."""
    return Planes(self.hi[i * B:(i + 1) * B], None if self.lo is None else self.lo[i * B:(i + 1) * B], self.scale)


class ConvLayer(object):
  """One conv (+BN)(+ReLU)(+x2 upsample) block: IMMModel.conv -> nnu.conv_block (nn_utils.py:151-210)."""

  def __init__(self, eng, prefix, name, k, stride, cin, cout, N, H, W, xcs, bn, relu, up2x, needs_dgrad,
               trainable=True, epilogue=_lib.EPI_BIAS):
    self.eng, self.prefix, self.name = eng, prefix, name
    self.k, self.stride, self.cin, self.cout = k, stride, cin, cout
    self.N, self.H, self.W, self.xcs = N, H, W, xcs
    self.Ho, self.Wo = -(-H // stride), -(-W // stride)
    self.bn, self.relu, self.up2x, self.needs_dgrad, self.trainable = bn, relu, up2x, needs_dgrad, trainable
    self.cin_pad = round_up(cin, 32)
    self.ycs = round_up(cout, 4)          # channel stride of y / dy (TMA needs 16-byte pixel strides; 8 channels for fp16 planes)
    self.h16 = False                      # operands (x, dy, packed weights) are scaled fp16 planes (IMMEngine._pick_formats)
    self.w_scale = self.w_amax = None
    self.pad_t = same_pad(H, k, stride)[0]
    self.pad_l = same_pad(W, k, stride)[0]
    self.epilogue = epilogue
    self.engine_override = None
    self.precision_override = None
    self.x_layout = _lib.XLAYOUT_NHWC

  def precision(self):
    if self.precision_override is not None:
      return self.precision_override
    if self.eng.h16 and not self.h16:
      return _lib.PREC_TF32X3               # a layer without fp16-plane kernels inside an fp16 engine
    return self.eng.precision

  def desc(self, N=None, x=None, y=None):
    """x / y: the Planes this call reads as input activations / touches as output-side planes (y of a forward call
    that writes planes, dy of dgrad / wgrad); their scale records travel in the descriptor (H16 layers only)."""
    d = ConvDesc()
    d.N, d.H, d.W, d.Cin = (self.N if N is None else N), self.H, self.W, self.cin
    d.Cout, d.kh, d.kw, d.stride = self.cout, self.k, self.k, self.stride
    d.Ho, d.Wo, d.pad_t, d.pad_l = self.Ho, self.Wo, self.pad_t, self.pad_l
    d.x_cstride, d.y_cstride, d.cin_pad = self.xcs, self.ycs, self.cin_pad
    d.epilogue = self.epilogue
    d.precision = self.precision()
    d.engine = self.eng.engine if self.engine_override is None else self.engine_override
    d.x_layout = self.x_layout
    if self.h16:
      d.w_scale = self.w_scale.data_ptr()
      if x is not None and x.scale is not None:
        d.x_scale = x.scale.data_ptr()
      if y is not None and y.scale is not None:
        d.y_scale = y.scale.data_ptr()
    return d

  def engines(self):
    d = self.desc()
    return tuple(_lib.lib().immb_conv_engine_for(d, op) for op in range(3))


class IMMEngine(object):
  """Buffers + kernel schedule for one model replica on one GPU.

  config: attribute-style object with the reference's `model` section keys (configs/experiments/*.yaml):
  n_maps, n_filters, n_filters_render, gauss_std, gauss_mode, renderer_stride, min_res, loss_mask,
  channels_bug_fix, perceptual.comp, reconstruction_loss, perceptual.l2."""

  def __init__(self, config, batch, image_size=128, device='cuda:0', precision=None,
               engine=_lib.ENGINE_AUTO, world_size=1, vgg_tf32_weights=True, streams=None, use_graph=None):
    if not torch.cuda.is_available():
      raise _lib.ImmbError('IMMEngine needs a CUDA device; there is no CPU fallback')
    _lib.lib()
    self.cfg = config
    self.B, self.R = int(batch), int(image_size)
    self.dev = torch.device(device)
    if precision is None:
      # default: scaled fp16 split operands on the tensor cores (IMMB_PREC_F16X3; every layer without an fp16-plane
      # kernel stays on 3xTF32); IMMB_PRECISION=tf32x3 selects the all-TF32 engine.  The SIMT cross-check engine reads
      # fp32 planes only.
      name = os.environ.get('IMMB_PRECISION', 'f16x3').lower()
      precision = {'f16x3': _lib.PREC_F16X3, 'tf32x3': _lib.PREC_TF32X3, 'tf32': _lib.PREC_TF32}[name]
      if engine == _lib.ENGINE_SIMT:
        precision = _lib.PREC_TF32X3
    if precision == _lib.PREC_F16X3 and engine == _lib.ENGINE_SIMT:
      raise _lib.ImmbError('the SIMT engine reads fp32 planes: use a TF32 precision with it')
    self.precision, self.engine = precision, engine
    self.h16 = precision == _lib.PREC_F16X3
    self.world_size = world_size
    # The frozen VGG16 weights are rounded to TF32 (round-to-nearest-even) once at load: their lo plane is then
    # exactly zero and the tower runs the 2-pass IMMB_PREC_TF32X2 product (hi*w + lo*w) instead of 3 passes.
    # Effect of the rounding, measured against the fp64 oracle with exact weights: loss 1e-7, level losses <= 1.6e-4,
    # weight gradients 1.7e-4 (median) -- below fp32's own rounding noise on this problem.
    self.vgg_tf32_weights = bool(vgg_tf32_weights) and precision in (_lib.PREC_TF32X3, _lib.PREC_F16X3)
    self.K = int(config.n_maps)
    if config.gauss_mode != 'rot':
      raise _lib.ImmbError("only gauss_mode 'rot' (used by every shipped config) is built; got %r" % config.gauss_mode)
    if config.reconstruction_loss != 'perceptual' or not config.perceptual.l2:
      raise _lib.ImmbError('only reconstruction_loss=perceptual with perceptual.l2=True is built')
    if int(config.renderer_stride) != 2 or int(config.min_res) != 16:
      raise _lib.ImmbError('renderer_stride 2 / min_res 16 expected')
    self.comp = list(config.perceptual.comp)
    self.use_mask = bool(config.loss_mask)
    self.n_extra = len(self.comp) if getattr(config, 'channels_bug_fix', False) else 0
    self.n_out = 3 + self.n_extra
    self.nf, self.nfr = int(config.n_filters), int(config.n_filters_render)
    self.inv_std = 1.0 / float(config.gauss_std)
    self.global_step = -1.0           # scripts/train.py:87-89,192: initialised to --reset-global-step (default -1)
    self.adam_t = 0
    self.adam_betas = (0.9, 0.999)
    # H16 scale records (include/imm_b200.h "H16 planes"): one int32 {e, amax bits} per fp16 tensor
    self.scale_recs = torch.zeros((768, 2), dtype=torch.int32, device=self.dev)
    self.n_scale_recs = 0
    self.h16_overflow = torch.zeros(1, dtype=torch.int32, device=self.dev)
    self._scale_mode, self._mode_scales, self._calibrating = None, {}, False
    self._build_layers()
    self._pick_formats()
    self._alloc_params()
    self._alloc_buffers()
    self.vgg_loaded = False
    # Stream-level concurrency of INDEPENDENT work (no arithmetic changes, same kernels, same results):
    #   bit 0: every weight-gradient conv runs on a side stream (dw is only needed by the optimiser), so the HBM-bound
    #          BN-backward kernels of the next layer overlap the tensor-bound wgrad kernels;
    #   bit 1: the pose-encoder branch (forward and backward) runs on its own stream next to the image-encoder branch;
    #   bit 2: the ground-truth half of the frozen VGG16 tower (features of future_image, which depend on the input
    #          batch alone) runs on its own stream next to the encoders / renderer; the predicted half follows the renderer.
    if streams is None:
      streams = int(os.environ.get('IMMB_STREAMS', '3'))   # bit 2 measured: no gain at batch 64 (20.19 vs 20.15 ms) -> off by default
    self.streams = int(streams)
    self.wgrad_stream = torch.cuda.Stream(device=self.dev) if self.streams & 1 else None
    self.pose_stream = torch.cuda.Stream(device=self.dev) if self.streams & 2 else None
    self.gt_stream = torch.cuda.Stream(device=self.dev) if self.streams & 4 else None
    # CUDA-graph replay of the training step (train_step only; forward / backward / optimizer_step stay eager)
    self.use_graph = bool(int(os.environ.get('IMMB_GRAPH', '1'))) if use_graph is None else bool(use_graph)
    self.fuse_bn_stats = bool(int(os.environ.get('IMMB_FUSE_BN_STATS', '1')))
    self.fuse_level_sums = bool(int(os.environ.get('IMMB_FUSE_LEVEL_SUMS', '1')))
    self._graphs, self._graph_key, self._graph_warm = None, None, 0
    # N > 1: bucketed all-reduce overlapped with the encoders' backward (IMMB_AR_OVERLAP=0: one all-reduce after backward)
    # and captured into the step graph (IMMB_GRAPH_NCCL=0: eager all-reduce between a fwd+bwd graph and an optimiser graph)
    self.overlap_allreduce = bool(int(os.environ.get('IMMB_AR_OVERLAP', '1')))
    self.graph_nccl = bool(int(os.environ.get('IMMB_GRAPH_NCCL', '1')))
    # graphs that captured NCCL kernels must be gone before the communicator is torn down (see release_graphs)
    import atexit
    import weakref
    ref = weakref.ref(self)
    atexit.register(lambda: ref() is not None and ref().release_graphs())
    self.comm_stream, self._allreduce_fn = None, None
    self._pack_table = None
    self.graph_replays, self.graph_launches_per_step = 0, 0
    self._events = {}

  # ------------------------------------------------------------------------------------------------
  # construction
  # ------------------------------------------------------------------------------------------------
  def _build_layers(self):
    B, R, K = self.B, self.R, self.K
    self.enc_feat = 8 * self.nf
    self.Cj = round_up(self.enc_feat + K, 32)          # channel stride of the renderer's concat input
    self.layers = OrderedDict()                        # key = TF scope prefix of the conv

    def add(prefix, name, k, stride, cin, cout, H, W, xcs, bn, relu, up2x, needs_dgrad):
      L = ConvLayer(self, prefix, name, k, stride, cin, cout, B, H, W, xcs, bn, relu, up2x, needs_dgrad)
      self.layers['%s/%s' % (prefix, name)] = L
      return L

    self.enc_layers = {}
    for enc in ('image_encoder', 'pose_encoder'):
      prefix = 'model/%s/encoder' % enc
      cin, size, lst = 3, R, []
      for i, (name, k, stride, cout) in enumerate(encoder_spec(self.nf)):
        L = add(prefix, name, k, stride, cin, cout, size, size, cin, True, True, False, i > 0)
        if i == 0 and self.engine != _lib.ENGINE_SIMT and k == 7 and size % 16 == 0:
          # first layer (7x7, Cin=3) runs on the tensor cores from a staged row-window image (include/imm_b200.h)
          L.x_layout, L.xcs = _lib.XLAYOUT_ROWWIN4, 4
        lst.append(L)
        cin, size = cout, L.Ho
      self.enc_layers[enc] = lst
    self.enc_out_size = size                           # 16 for R=128, 32 for R=256
    self.pose_conv = add('model/pose_encoder', 'conv_1', 1, 1, self.enc_feat, K, size, size, self.enc_feat,
                         False, False, False, True)
    self.ren_layers = []
    cin, size, xcs = self.enc_feat + K, 16, self.Cj
    for name, cout, bn, relu, up in renderer_spec(self.nfr, R, self.n_out):
      L = add('model/renderer', name, 3, 1, cin, cout, size, size, xcs, bn, relu, up, True)
      self.ren_layers.append(L)
      cin, xcs = cout, cout
      if up:
        size *= 2
    # producer of each conv's input when that producer's BN-backward sums can ride in this conv's dgrad epilogue
    for lst in list(self.enc_layers.values()) + [self.ren_layers]:
      for i, L in enumerate(lst):
        P = lst[i - 1] if i > 0 else None
        L.producer = P if (P is not None and P.bn and not P.up2x and P.cout == L.cin and L.xcs == L.cin) else None
    self.pose_conv.producer = None
    # frozen VGG16 up to the deepest level the loss reads (conv3_3 / conv4_3 feed pools)
    needed = [c for c in self.comp if c != 'input']
    last = max((i for i, it in enumerate(VGG_ORDER) if not isinstance(it, str) and it[0] in needed), default=-1)
    self.vgg_seq = []
    cin, size = 1, R
    for it in VGG_ORDER[:last + 1]:
      if isinstance(it, str):
        self.vgg_seq.append(('pool', it, cin, size))
        size //= 2
      else:
        name, cout = it
        if name == 'conv1_1':
          # Cin = 1, 3x3 == 1x1 convolution over the [.,R,R,12] tensor of 3x3 patches written by immb_vgg_prologue
          L = ConvLayer(self, 'SelfSupReconstructionLoss/vgg16', name, 1, 1, 9, cout, 2 * B, size, size, 12,
                        False, True, False, True, trainable=False, epilogue=_lib.EPI_BIAS_RELU)
        else:
          L = ConvLayer(self, 'SelfSupReconstructionLoss/vgg16', name, 3, 1, cin, cout, 2 * B, size, size, cin,
                        False, True, False, True, trainable=False, epilogue=_lib.EPI_BIAS_RELU)
        self.vgg_seq.append(('conv', L, cin, size))
        cin = cout

  def _new_scale(self, enable=True):
    """A fresh scale record (view of two int32) or None."""
    if not enable:
      return None
    i = self.n_scale_recs
    self.n_scale_recs += 1
    assert i < self.scale_recs.shape[0]
    if not hasattr(self, 'scale_tags'):
      self.scale_tags = []
    import sys
    fr = sys._getframe(1)
    self.scale_tags.append('%s:%d' % (fr.f_code.co_name, fr.f_lineno))      # who owns record i (diagnostics: tools/soak_step.py)
    return self.scale_recs[i]

  def _vgg_convs(self):
    return [item for kind, item, cin, size in self.vgg_seq if kind == 'conv']

  def _pick_formats(self):
    """Which layers run on scaled fp16 planes: every conv whose forward, dgrad (if needed) and wgrad (if trainable) all
    have an fp16-plane tensor-core kernel (the stride-1 and stride-2 3x3 layers and the frozen tower); the 7x7 first
    layers and the 1x1 heat-map conv stay on 3xTF32.  The planes BETWEEN two layers take the consumer's format."""
    if not self.h16:
      return
    lib = _lib.lib()

    def eligible(L):
      if L.x_layout != _lib.XLAYOUT_NHWC or L.xcs % 8:
        return False
      d = L.desc()
      d.precision, d.y_cstride = _lib.PREC_F16X3, round_up(L.cout, 8)
      ops = [0] + ([1] if L.needs_dgrad else []) + ([2] if L.trainable else [])
      return all(lib.immb_conv_engine_for(d, op) == _lib.ENGINE_TC for op in ops)

    for L in list(self.layers.values()) + self._vgg_convs():
      if L.name == 'conv1_1' and not L.trainable:
        continue
      if eligible(L):
        L.h16, L.ycs = True, round_up(L.cout, 8)
    convs = self._vgg_convs()
    if len(convs) > 1 and convs[0].name == 'conv1_1':
      convs[0].h16 = convs[1].h16            # conv1_1 has its own CUDA-core kernels: only its planes follow conv1_2

  def _alloc_params(self):
    dev = self.dev
    names, shapes, wds = [], [], []

    def reg(name, shape, wd=0.0):
      names.append(name)
      shapes.append(tuple(shape))
      wds.append(wd)

    for key, L in self.layers.items():
      reg('%s/%s/w' % (key, L.name), (L.k, L.k, L.cin, L.cout), WD)     # nn_utils.py:44-46 (w only)
      reg('%s/%s/b' % (key, L.name), (L.cout,))
      if L.bn:
        reg('%s/batch_normalization/gamma' % key, (L.cout,))
        reg('%s/batch_normalization/beta' % key, (L.cout,))
    self.param_names, self.param_shapes = names, shapes
    offs, off = [], 0
    for s in shapes:
      offs.append(off)
      off += round_up(int(np.prod(s)), 4)
    self.param_offsets, self.n_flat = offs, off
    self.ren_grad_offset = offs[names.index('model/renderer/conv_1/conv_1/w')]      # renderer tensors are the tail of the buffer
    assert all(n.startswith('model/renderer/') for n in names[names.index('model/renderer/conv_1/conv_1/w'):])
    z = lambda: torch.zeros(off, dtype=torch.float32, device=dev)
    self.flat_p, self.flat_g, self.flat_m, self.flat_v = z(), z(), z(), z()
    view = lambda flat: OrderedDict((n, flat[o:o + int(np.prod(s))].view(s)) for n, o, s in zip(names, offs, shapes))
    self.params, self.grads = view(self.flat_p), view(self.flat_g)
    self.adam_m, self.adam_v = view(self.flat_m), view(self.flat_v)
    # optimiser chunk table
    ct, co, cl = [], [], []
    for t, (o, s) in enumerate(zip(offs, shapes)):
      n = int(np.prod(s))
      for c0 in range(0, n, ADAM_CHUNK):
        ct.append(t)
        co.append(o + c0)
        cl.append(min(ADAM_CHUNK, n - c0))
    self.chunk_tensor = torch.tensor(ct, dtype=torch.int32, device=dev)
    self.chunk_off = torch.tensor(co, dtype=torch.int64, device=dev)
    self.chunk_len = torch.tensor(cl, dtype=torch.int32, device=dev)
    self.n_chunks = len(ct)
    self.tensor_wd = torch.tensor(wds, dtype=torch.float32, device=dev)
    self.n_tensors = len(names)
    self.sq = torch.zeros(2 * self.n_tensors, dtype=torch.float64, device=dev)     # [sq ; wsq]
    self.w_amax = torch.zeros(self.n_tensors, dtype=torch.float32, device=dev)     # per-tensor max |p| (fp16 weight scaling)
    # non-trainable state: BN moving stats + loss normalisers
    bnames, bshapes = [], []
    for key, L in self.layers.items():
      if L.bn:
        bnames += ['%s/batch_normalization/moving_mean' % key, '%s/batch_normalization/moving_variance' % key]
        bshapes += [(L.cout,), (L.cout,)]
    boffs, off = [], 0
    for s in bshapes:
      boffs.append(off)
      off += s[0]
    self.flat_bn = torch.zeros(off, dtype=torch.float32, device=dev)
    self.buffers = OrderedDict((n, self.flat_bn[o:o + s[0]]) for n, o, s in zip(bnames, boffs, bshapes))
    for n in bnames:
      if n.endswith('moving_variance'):
        self.buffers[n].fill_(1.0)
    self.agg = torch.tensor(PERCEPTUAL_WS[:len(self.comp)], dtype=torch.float32, device=dev)
    for k, nm in enumerate(self.comp):
      self.buffers['SelfSupReconstructionLoss/%s_agg' % nm] = self.agg[k:k + 1].view(())
    # per-layer handles
    for key, L in self.layers.items():
      L.w, L.b = self.params['%s/%s/w' % (key, L.name)], self.params['%s/%s/b' % (key, L.name)]
      t = names.index('%s/%s/w' % (key, L.name))
      L.w_amax = self.w_amax[t:t + 1]
      L.dw, L.db = self.grads['%s/%s/w' % (key, L.name)], self.grads['%s/%s/b' % (key, L.name)]
      if L.bn:
        pre = '%s/batch_normalization/' % key
        L.gamma, L.beta = self.params[pre + 'gamma'], self.params[pre + 'beta']
        L.dgamma, L.dbeta = self.grads[pre + 'gamma'], self.grads[pre + 'beta']
        L.mm, L.mv = self.buffers[pre + 'moving_mean'], self.buffers[pre + 'moving_variance']

  def _alloc_weight_planes(self, L):
    dev, taps = self.dev, L.k * L.k
    e = torch.empty
    if L.x_layout == _lib.XLAYOUT_ROWWIN4:
      L.wp = Planes(e((7, L.cout, 32), dtype=torch.float32, device=dev), e((7, L.cout, 32), dtype=torch.float32, device=dev))
      L.wh = Planes(None, None)
      L.stage = Planes.alloc((L.N, L.H, L.W + 8, 4), dev)
      return
    dt = torch.float16 if L.h16 else torch.float32
    L.w_scale = self._new_scale(L.h16)
    L.wp = Planes(e((taps, L.cout, L.cin_pad), dtype=dt, device=dev), e((taps, L.cout, L.cin_pad), dtype=dt, device=dev))
    L.wh = Planes(e((taps, L.cin_pad, L.ycs), dtype=dt, device=dev), e((taps, L.cin_pad, L.ycs), dtype=dt, device=dev))

  def _alloc_buffers(self):
    dev, B, R, K = self.dev, self.B, self.R, self.K
    f32 = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
    # double accumulators live in two pools that are cleared with ONE memset per pass (fwd / bwd)
    n_fwd = sum(2 * L.cout for L in self.layers.values() if L.bn) + len(self.comp)
    n_bwd = sum((2 * L.cout if L.bn else 0) + L.cout for L in self.layers.values())
    self.fwd_pool = torch.zeros(n_fwd, dtype=torch.float64, device=dev)
    self.bwd_pool = torch.zeros(n_bwd, dtype=torch.float64, device=dev)
    cur = {'f': 0, 'b': 0}

    def take(pool, key, n):
      v = pool[cur[key]:cur[key] + n]
      cur[key] += n
      return v
    f64z = lambda n: torch.zeros(n, dtype=torch.float64, device=dev)
    ns = self._new_scale
    self.joint = Planes.alloc((B, 16, 16, self.Cj), dev, zero=True, scale=ns(self.ren_layers[0].h16))     # pad channels stay zero
    ws_bytes = 0
    for key, L in self.layers.items():
      self._alloc_weight_planes(L)
      L.y = torch.zeros((B, L.Ho, L.Wo, L.ycs), dtype=torch.float32, device=dev)
      L.dy = Planes.alloc((B, L.Ho, L.Wo, L.ycs), dev, zero=True, scale=ns(L.h16))
      L.dbias_acc = take(self.bwd_pool, 'b', L.cout)
      if L.bn:
        L.sums, L.bsums = take(self.fwd_pool, 'f', 2 * L.cout), take(self.bwd_pool, 'b', 2 * L.cout)
        L.scale, L.shift, L.mean, L.invstd = f32(L.cout), f32(L.cout), f32(L.cout), f32(L.cout)
        L.g_low = f32(B, L.Ho, L.Wo, L.cout) if L.up2x else None
      if L.needs_dgrad:
        L.dx = torch.zeros((B, L.H, L.W, L.xcs), dtype=torch.float32, device=dev)
      ws_bytes = max(ws_bytes, int(call('immb_conv2d_wgrad_workspace', L.desc())))
    # activation planes: output of each trainable block
    # activation planes take the CONSUMER's format
    for enc in ('image_encoder', 'pose_encoder'):
      lst = self.enc_layers[enc]
      for i, L in enumerate(lst):
        last = i == len(lst) - 1
        if enc == 'image_encoder' and last and L.Ho == 16:
          L.out, L.ocs = self.joint, self.Cj               # written straight into the concat buffer
        else:
          consumer = lst[i + 1] if not last else (self.pose_conv if enc == 'pose_encoder' else self.ren_layers[0])
          L.out, L.ocs = Planes.alloc((B, L.Ho, L.Wo, L.cout), dev, scale=ns(consumer.h16)), L.cout
    if self.enc_out_size != 16:
      self.g_enc_resized = f32(B, self.enc_out_size, self.enc_out_size, self.enc_feat)
    for i, L in enumerate(self.ren_layers):
      if L.bn:
        s = 2 if L.up2x else 1
        L.out, L.ocs = Planes.alloc((B, L.Ho * s, L.Wo * s, L.cout), dev, scale=ns(self.ren_layers[i + 1].h16)), L.cout
    S = self.enc_out_size
    self.mu, self.py, self.px = f32(B, K, 2), f32(B, S, K), f32(B, S, K)
    self.Kp = self.pose_conv.ycs
    self.g_heat = torch.zeros((B, S, S, self.Kp), dtype=torch.float32, device=dev)
    # perceptual tower
    self.vgg_in = Planes.alloc((2 * B, R, R, 12), dev)        # 3x3 patches of the normalised gray image (SIMT engine only)
    self.vgg_act = OrderedDict()
    for idx, (kind, item, cin, size) in enumerate(self.vgg_seq):
      later = [it for k, it, _, _ in self.vgg_seq[idx + 1:] if k == 'conv']
      if kind == 'conv':
        L = item
        self._alloc_weight_planes(L)
        L.w = f32(L.k, L.k, L.cin, L.cout)
        L.b = f32(L.cout)
        L.w_amax = torch.zeros(1, dtype=torch.float32, device=dev)
        consumer = later[0] if later else L
        L.out = Planes.alloc((2 * B, size, size, L.cout), dev, scale=ns(consumer.h16))
        L.dy = Planes.alloc((B, size, size, L.cout), dev, scale=ns(L.h16))        # pred half only
        L.dx = torch.zeros((B, size, size, L.xcs), dtype=torch.float32, device=dev)
        self.vgg_act[L.name] = L.out
        ws_bytes = max(ws_bytes, 0)
      else:
        fmt = later[0].h16 if later else False
        self.vgg_act[item] = Planes.alloc((2 * B, size // 2, size // 2, cin), dev, scale=ns(fmt))
    self.g_pool = {item: f32(B, size, size, cin) for kind, item, cin, size in self.vgg_seq if kind == 'pool'}
    self.level_acc = take(self.fwd_pool, 'f', len(self.comp))
    counts = []
    for nm in self.comp:
      if nm == 'input':
        counts.append(float(B * R * R * 3))
      else:
        P = self.vgg_act[nm].hi
        counts.append(float(B * P.shape[1] * P.shape[2] * P.shape[3]))
    self.level_counts = torch.tensor(counts, dtype=torch.float64, device=dev)
    self.levels = f32(len(self.comp))
    self.coef = f32(len(self.comp))
    self.rec_loss, self.weights_loss, self.total_loss = f32(1), f32(1), f32(1)
    self.pcs = self.ren_layers[-1].ycs      # channel stride of the renderer output / its gradient
    self.pred_dy = Planes.alloc((B, R, R, self.pcs), dev, zero=True, scale=ns(self.ren_layers[-1].h16))
    self.workspace = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    # scratch of the two-level (deterministic, atomic-free) BN reductions
    n_scr = max([int(call('immb_bn_scratch_elems', L.N * L.Ho * L.Wo, L.cout)) for L in self.layers.values() if L.bn] + [16])
    for L in self.layers.values():
      L.stats_rows = int(_lib.lib().immb_conv2d_fwd_stats_rows(L.desc())) if (L.bn and L.ycs == L.cout) else 0
      n_scr = max(n_scr, L.stats_rows * 2 * L.cout)
      L.bwd_rows = 0                   # > 0 while this layer's BN-backward partials sit in the scratch buffer
      L.dgrad_stats_rows = (int(_lib.lib().immb_conv2d_dgrad_stats_rows(L.desc()))
                            if (L.needs_dgrad and getattr(L, 'producer', None) is not None) else 0)
      n_scr = max(n_scr, L.dgrad_stats_rows * 2 * L.cin)
    # deferred bias-gradient reduction: per-layer partial buffers + device tables of immb_reduce_item
    import ctypes
    self._db_table = None
    items, ren_items = [], []
    for key, L in self.layers.items():
      L.db_partials = None
      if not L.bn:
        continue
      nb = int(call('immb_bn_bwd_apply_blocks', L.N * L.Ho * L.Wo, L.cout))
      if nb <= 0:
        continue
      L.db_partials = torch.empty(nb * L.cout, dtype=torch.float64, device=dev)
      it = (L.db_partials.data_ptr(), L.db.data_ptr(), nb, L.cout)
      (ren_items if key.startswith('model/renderer/') else items).append(it)

    def table(lst):
      arr, b0 = (_lib.ReduceItem * len(lst))(), 0
      for i, (part, out, nb, nv) in enumerate(lst):
        arr[i].partials, arr[i].out_f, arr[i].nblocks, arr[i].nvals, arr[i].block0 = part, out, nb, nv, b0
        b0 += (nv + 31) // 32
      raw = torch.frombuffer(bytearray(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr))), dtype=torch.uint8)
      return raw.to(dev), len(lst), b0
    if items and ren_items:
      self._db_table_enc, self._db_n_enc, self._db_blocks_enc = table(items)
      self._db_table_ren, self._db_n_ren, self._db_blocks_ren = table(ren_items)
      self._db_table, self._db_n, self._db_blocks = table(items + ren_items)
    self.bn_scratch = torch.empty(n_scr, dtype=torch.float64, device=dev)
    self.bn_scratch_pose = torch.empty(n_scr, dtype=torch.float64, device=dev)      # the pose-branch stream's own scratch

  # ------------------------------------------------------------------------------------------------
  # parameters
  # ------------------------------------------------------------------------------------------------
  def init_parameters(self, seed=0):
    """Reference initialisers: w ~ truncated_normal(std .01) (nn_utils.py:47), b = 0 (:106), gamma 1,
    beta 0, moving_mean 0, moving_var 1 (tf.layers defaults), *_agg = ws (imm_model.py:131)."""
    gen = torch.Generator().manual_seed(seed)
    for key, L in self.layers.items():
      w = torch.empty(L.w.shape, dtype=torch.float32)
      torch.nn.init.trunc_normal_(w, mean=0.0, std=INIT_STD, a=-2 * INIT_STD, b=2 * INIT_STD, generator=gen)
      L.w.copy_(w)
      L.b.zero_()
      if L.bn:
        L.gamma.fill_(1.0)
        L.beta.zero_()
        L.mm.zero_()
        L.mv.fill_(1.0)
    self.agg.copy_(torch.tensor(PERCEPTUAL_WS[:len(self.comp)]))
    self.flat_m.zero_()
    self.flat_v.zero_()
    self.adam_t = 0
    self.adam_betas = (0.9, 0.999)
    self.global_step = -1.0
    self._invalidate_scales()
    self.repack_weights()

  def load_state(self, params=None, buffers=None, adam_m=None, adam_v=None):
    """Copies tensors keyed by TF variable name (HWIO weights) into the device buffers."""
    for src, dst in ((params, self.params), (buffers, self.buffers), (adam_m, self.adam_m), (adam_v, self.adam_v)):
      if src is None:
        continue
      for k, v in src.items():
        if k in dst:
          dst[k].copy_(torch.as_tensor(v, dtype=torch.float32).reshape(dst[k].shape))
    self._invalidate_scales()
    self.repack_weights()

  def load_vgg_caffe_dict(self, data):
    """vgg16.py:17-47,74-92 with pre_adjust_batch_norm=True (build_vgg16.py:30): OIHW->HWIO, BGR flip only for
    Cin==3, BN folding W /= sigma, b = (b-mu)/sigma, sigma = sqrt(1e-5 + bn['1']/bn['2']), mu = bn['0']/bn['2']."""
    self.vgg_params = OrderedDict()
    for kind, item, cin, size in self.vgg_seq:
      if kind != 'conv':
        continue
      L = item
      W = np.array(data[L.name]['0'], dtype=np.float32).copy().transpose(2, 3, 1, 0)
      if L.name == 'conv1_1' and W.shape[2] == 3:
        W = W[:, :, ::-1]
      bias = np.array(data[L.name]['1'], dtype=np.float32).copy()
      bn_name = 'batch_' + L.name
      if bn_name in data:
        bn = data[bn_name]
        sigma = np.sqrt(1e-5 + np.asarray(bn['1']) / np.asarray(bn['2']))
        mu = np.asarray(bn['0']) / np.asarray(bn['2'])
        W = W / sigma
        bias = (bias - mu) / sigma
      self._set_vgg_layer(L, W, bias)
    self.vgg_loaded = True

  def load_vgg_hwio(self, variables):
    """Frozen tower from already-prepared TF variables {'SelfSupReconstructionLoss/vgg16/<conv>/weights': HWIO,
    '.../biases': [Cout]} -- the form the reference's checkpoints hold them in (tf.global_variables() are all saved,
    cnn_train_multi.py:439; vgg16.py:175-179), BN already folded.  Returns False if a needed layer is missing."""
    pre = 'SelfSupReconstructionLoss/vgg16/'
    convs = [item for kind, item, cin, size in self.vgg_seq if kind == 'conv']
    if not all((pre + L.name + '/weights') in variables and (pre + L.name + '/biases') in variables for L in convs):
      return False
    self.vgg_params = OrderedDict()
    for L in convs:
      W = np.ascontiguousarray(np.asarray(variables[pre + L.name + '/weights'], dtype=np.float32))
      bias = np.asarray(variables[pre + L.name + '/biases'], dtype=np.float32)
      self._set_vgg_layer(L, W, bias)
    self.vgg_loaded = True
    return True

  def _set_vgg_layer(self, L, W, bias):
    """W: HWIO float32 (BN folded), bias [Cout]."""
    cin_true = 1 if L.name == 'conv1_1' else L.cin
    assert tuple(W.shape) == (3, 3, cin_true, L.cout), 'Incorrect weights shape for %s' % L.name   # vgg16.py:171
    W = np.ascontiguousarray(W, dtype=np.float32)
    if self.vgg_tf32_weights and L.name != 'conv1_1':      # conv1_1 runs in exact fp32 on the CUDA cores
      if L.h16:
        # scaled fp16 planes: the pair kernel's 2-pass product reads the hi weight plane only, i.e. the weights rounded
        # to fp16's 11 significant bits (the same width as TF32) -- the rounding happens in immb_pack_weights
        L.precision_override = _lib.PREC_F16X2
      else:
        bits = W.view(np.uint32).astype(np.uint64)
        bits = (bits + 0x0FFF + ((bits >> 13) & 1)) & 0xFFFFE000          # round-to-nearest-even to 10 mantissa bits
        W = bits.astype(np.uint32).view(np.float32)
        L.precision_override = _lib.PREC_TF32X2
    L.w.copy_(torch.from_numpy(W).reshape(L.w.shape))
    L.w_amax.fill_(float(np.abs(W).max()))
    self._invalidate_scales()
    L.b.copy_(torch.from_numpy(np.ascontiguousarray(bias, dtype=np.float32)))
    self.vgg_params['SelfSupReconstructionLoss/vgg16/%s/weights' % L.name] = L.w.view(3, 3, cin_true, L.cout)
    self.vgg_params['SelfSupReconstructionLoss/vgg16/%s/biases' % L.name] = L.b
    self._pack(L)

  def _pack(self, L):
    if L.x_layout == _lib.XLAYOUT_ROWWIN4:
      call('immb_pack_weights_rowwin', L.w, L.cout, L.wp.hi, L.wp.lo, _lib.stream_ptr())
      return
    call('immb_pack_weights', L.w, L.k, L.k, L.cin, L.cout, L.cin_pad, L.ycs, L.wp.hi, L.wp.lo, L.wh.hi, L.wh.lo,
         L.w_amax if L.h16 else None, L.w_scale if L.h16 else None, _lib.stream_ptr())

  def repack_weights(self, amax_current=False):
    """Rebuilds the tensor-core weight planes from the master weights.  fp16 planes are scaled by each tensor's largest
    magnitude: the optimiser kernel leaves it in w_amax (amax_current), otherwise it is measured here."""
    if self.h16 and not amax_current:
      self.w_amax.zero_()
      call('immb_multi_amax', self.flat_p, self.chunk_tensor, self.chunk_off, self.chunk_len, self.n_chunks, self.w_amax,
           _lib.stream_ptr())
    if self._pack_table is None:
      self._build_pack_table()
    call('immb_pack_weights_multi', self._pack_table, self._pack_n, self._pack_blocks, _lib.stream_ptr())

  def _build_pack_table(self):
    """Device table of immb_pack_item records: every trainable weight tensor is repacked by ONE launch per step."""
    import ctypes
    items, block0 = [], 0
    ptr = lambda t: (t.data_ptr() if t is not None else None)
    for L in self.layers.values():
      it = _lib.PackItem()
      it.w = L.w.data_ptr()
      it.wp_hi, it.wp_lo = ptr(L.wp.hi), ptr(L.wp.lo)
      it.wh_hi, it.wh_lo = ptr(L.wh.hi), ptr(L.wh.lo)
      it.taps, it.Cin, it.Cout, it.cin_pad, it.cout_pad = L.k * L.k, L.cin, L.cout, L.cin_pad, L.ycs
      if L.x_layout == _lib.XLAYOUT_ROWWIN4:
        it.kind, n = 2, 7 * L.cout * 32
      else:
        it.kind, n = (1 if L.h16 else 0), L.k * L.k * L.cin_pad * L.ycs
        if L.h16:
          it.amax, it.rec = L.w_amax.data_ptr(), L.w_scale.data_ptr()
      it.block0, it.nblocks = block0, (n + 2047) // 2048
      block0 += it.nblocks
      items.append(it)
    arr = (_lib.PackItem * len(items))(*items)
    raw = torch.frombuffer(bytearray(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr))), dtype=torch.uint8)
    self._pack_table = raw.to(self.dev)
    self._pack_n, self._pack_blocks = len(items), block0

  # ------------------------------------------------------------------------------------------------
  # H16 scale records: delayed scaling + calibration
  # ------------------------------------------------------------------------------------------------
  def _invalidate_scales(self):
    self._scale_mode, self._mode_scales = None, {}

  def _scale_update(self):
    """Turns the maxima observed since the last update into the next exponents (include/imm_b200.h).  Runs at the START of
    a forward pass, when no fp16 activation / gradient tensor of the previous step is live any more."""
    if self.h16 and self.n_scale_recs:
      call('immb_scale_update', self.scale_recs, self.n_scale_recs, self.h16_overflow, _lib.stream_ptr())

  def _enter_scale_mode(self, image, future_image, mask, training, build_loss):
    """The exponents of the activation planes depend on what flows through the network: training-mode batch statistics
    vs inference-mode moving statistics, with or without a backward pass.  Each mode keeps its own set of exponents:
    switching modes (the periodic test pass of train_loop) stashes / restores them; a mode seen for the first time is
    calibrated by dry passes over the given inputs (model state is restored afterwards), repeated until the exponents
    stop moving, because a tensor written with a far-off exponent is only approximately right and so is everything
    computed from it."""
    mode = (bool(training), bool(build_loss))
    with_backward = mode == (True, True)          # a training forward with the loss is followed by backward()
    if self._scale_mode == mode or self._calibrating:
      return
    if self._scale_mode is not None:
      self._mode_scales[self._scale_mode] = self.scale_recs.clone()
    if mode in self._mode_scales:
      self.scale_recs.copy_(self._mode_scales[mode])
      self._scale_mode = mode
      return
    if (mode[0], True) in self._mode_scales:     # same statistics mode without the loss: a subset of a calibrated mode
      self.scale_recs.copy_(self._mode_scales[(mode[0], True)])
      self._scale_mode = mode
      return
    self._calibrating = True
    try:
      saved = (self.flat_bn.clone(), self.agg.clone())
      prev = None
      for it in range(8):
        self.forward(image, future_image, mask, training=training, build_loss=build_loss)
        if with_backward:
          self.backward()
        self._scale_update()
        cur = self.scale_recs[:self.n_scale_recs, 0].clone()
        self.flat_bn.copy_(saved[0])
        self.agg.copy_(saved[1])
        if prev is not None and int((cur - prev).abs().max().item()) <= 1:
          break
        prev = cur
      self.h16_overflow.zero_()
      self.calibration_passes = it + 1
    finally:
      self._calibrating = False
    self._scale_mode = mode

  # ------------------------------------------------------------------------------------------------
  # forward
  # ------------------------------------------------------------------------------------------------
  def _conv_fwd(self, L, X, y_hi, y_lo=None, N=None, Y=None):
    """Y: the output Planes when the result is written as split planes (their scale record goes into the descriptor)."""
    call('immb_conv2d_fwd', L.desc(N, x=X, y=Y), X.hi, X.lo, L.w, L.wp.hi, L.wp.lo, L.b, y_hi, y_lo, _lib.stream_ptr())

  def _event(self, key):
    ev = self._events.get(key)
    if ev is None:
      ev = self._events[key] = torch.cuda.Event()
    return ev

  def _fork(self, side, key):
    """side stream waits for everything enqueued so far on the current stream."""
    ev = self._event(key)
    ev.record()
    side.wait_event(ev)

  def _join(self, side, key):
    """current stream waits for everything enqueued so far on the side stream."""
    ev = self._event(key)
    ev.record(side)
    torch.cuda.current_stream().wait_event(ev)

  def _block_fwd(self, L, X, training, scratch=None):
    """conv -> bias -> [BN] -> [ReLU] -> [x2 legacy bilinear]  (nn_utils.py:151-210, imm_model.py:175)."""
    st = _lib.stream_ptr()
    _lib.TAG = 'fwd:%s/%s' % (L.prefix.split('/')[-2] if L.prefix.endswith('encoder') else L.prefix.split('/')[-1], L.name)
    L.x = X
    sc = self.bn_scratch if scratch is None else scratch
    fused_stats = bool(L.bn and training and self.fuse_bn_stats and L.stats_rows > 0)
    if fused_stats:
      # batch statistics accumulated in the conv epilogue (per-CTA partial rows -> fixed-order second level)
      call('immb_conv2d_fwd_bnstats', L.desc(x=X), X.hi, X.lo, L.wp.hi, L.wp.lo, L.b, L.y, sc, sc.numel(), st)
    else:
      self._conv_fwd(L, X, L.y)
    if not L.bn:
      return None
    npix = L.N * L.Ho * L.Wo
    if fused_stats:
      # second level of the epilogue statistics + finalize in one launch
      call('immb_bn_finalize_partials', sc, L.stats_rows, npix, L.cout, L.gamma, L.beta, L.mm, L.mv, L.scale, L.shift,
           L.mean, L.invstd, st)
    else:
      if training:
        call('immb_bn_stats', L.y, npix, L.cout, L.ycs, L.sums, sc, sc.numel(), st)
      call('immb_bn_finalize', L.sums, npix, L.cout, L.gamma, L.beta, L.mm, L.mv, 1 if training else 0,
           L.scale, L.shift, L.mean, L.invstd, st)
    call('immb_bn_apply', L.y, L.N, L.Ho, L.Wo, L.cout, L.ycs, L.scale, L.shift, 1 if L.relu else 0,
         1 if L.up2x else 0, L.out.hi, L.out.lo, L.ocs, L.out.scale, st)
    return L.out

  # sections of the forward pass (also the bodies of IMMModel's sub-builders image_encoder / pose_encoder / simple_renderer)
  def run_encoder(self, enc, inp, training, scratch=None):
    """IMMModel.encoder (imm_model.py:182-217) of branch `enc` on inp [B,R,R,3]."""
    B, R = self.B, self.R
    X = Planes(inp, None)
    L0 = self.enc_layers[enc][0]
    if L0.x_layout == _lib.XLAYOUT_ROWWIN4:
      call('immb_stage_image_rowwin', inp, B, R, R, L0.stage.hi, L0.stage.lo, _lib.stream_ptr())
      X = L0.stage
    for L in self.enc_layers[enc]:
      X = self._block_fwd(L, X, training, scratch)
    return X

  def run_image_encoder(self, image, training):
    """image_encoder (imm_model.py:220-230); the 16x16 block lands in channels [0, enc_feat) of the concat buffer
    (resize_bilinear(align_corners=True) first when the encoder ends above the render size, :324-335)."""
    self.run_encoder('image_encoder', image, training, None)
    img_last = self.enc_layers['image_encoder'][-1]
    if self.enc_out_size != 16:
      S = self.enc_out_size
      call('immb_resize_ac_fwd', img_last.out.hi, img_last.out.lo, img_last.cout, self.B, S, S, self.enc_feat, 16, 16,
           self.joint.hi, self.joint.lo, self.Cj, img_last.out.scale, self.joint.scale, _lib.stream_ptr())

  def run_pose_branch(self, future_image, training, scratch=None):
    """pose encoder -> heatmaps -> (mu_y, mu_x) -> Gaussian maps into channels [enc_feat, enc_feat+K) of the concat
    buffer (imm_model.py:233-274,341-344)."""
    self.run_encoder('pose_encoder', future_image, training, scratch)
    pose_last = self.enc_layers['pose_encoder'][-1]
    self._block_fwd(self.pose_conv, pose_last.out, training, scratch)
    S = self.enc_out_size
    call('immb_softargmax_gauss_fwd', self.pose_conv.y, self.B, S, self.K, self.Kp, self.inv_std, self.mu, self.py,
         self.px, 16, self.joint.hi, self.joint.lo, self.Cj, self.enc_feat, self.joint.scale, _lib.stream_ptr())

  def run_renderer(self, training):
    """simple_renderer (imm_model.py:154-179) on the concat buffer -> self.pred [B,R,R,pcs]."""
    X = self.joint
    for L in self.ren_layers:
      X = self._block_fwd(L, X, training)
    self.pred = self.ren_layers[-1].y
    return self.pred

  def load_joint(self, feat16):
    """Fills the renderer's concat buffer from an explicit [B,16,16,enc_feat+K] tensor (simple_renderer called on its own)."""
    B, C = self.B, self.enc_feat + self.K
    assert tuple(feat16.shape) == (B, 16, 16, C), 'renderer input must be [B,16,16,%d]' % C
    padded = torch.zeros((B, 16, 16, self.Cj), dtype=torch.float32, device=self.dev)
    padded[..., :C] = feat16.to(self.dev, torch.float32)
    call('immb_split_planes', padded, self.joint.hi, self.joint.lo, padded.numel(), self.joint.scale, _lib.stream_ptr())

  def forward(self, image, future_image, mask=None, training=True, build_loss=True):
    """IMMModel.build (imm_model.py:413-490).  image / future_image [B,R,R,3] fp32 in [0,255]; mask [B,R,R,1]."""
    st = _lib.stream_ptr()
    B, R, K = self.B, self.R, self.K
    assert tuple(image.shape) == (B, R, R, 3) and tuple(future_image.shape) == (B, R, R, 3)
    self.image, self.future_image = image, future_image
    self.mask = mask if (self.use_mask and mask is not None) else None
    if self.use_mask and mask is None and build_loss:
      raise RuntimeError('No loss mask recieved but is required.')      # imm_model.py:363-367
    self.training = training
    if self.h16:
      self._enter_scale_mode(image, future_image, mask, training, build_loss)
      self._scale_update()
    self.fwd_pool.zero_()
    self._gt_tower_forked = False
    if self.gt_stream is not None and build_loss and self.vgg_loaded and self.engine != _lib.ENGINE_SIMT:
      self._fork(self.gt_stream, 'gt_fork')
      with torch.cuda.stream(self.gt_stream):
        self._vgg_tower(0)
      self._gt_tower_forked = True
    # image encoder (imm_model.py:220-230) next to the pose branch (:233-274), then the renderer (:154-179)
    if self.pose_stream is not None:
      self._fork(self.pose_stream, 'fwd_fork')
      with torch.cuda.stream(self.pose_stream):
        self.run_pose_branch(future_image, training, self.bn_scratch_pose)
    self.run_image_encoder(image, training)
    if self.pose_stream is not None:
      self._join(self.pose_stream, 'fwd_join')
    else:
      self.run_pose_branch(future_image, training, None)
    self.run_renderer(training)
    self.pred = self.ren_layers[-1].y           # [B,R,R,pcs]; first 3 channels = future_im_pred (:348-355)
    if build_loss:
      self._loss_fwd(training)
    return self.pred

  def _vgg_tower(self, which):
    """build_vgg16 (build_vgg16.py:14-35) + vgg16.build_network (vgg16.py:289-375) on [future_image ; pred].
    which: None = both halves as one batch of 2B, 0 = ground-truth half only, 1 = predicted half only."""
    st = _lib.stream_ptr()
    B, R = self.B, self.R
    n = 2 * B if which is None else B
    sel = (lambda P: P) if which is None else (lambda P: P.half(which, B))
    fused_first = self.engine != _lib.ENGINE_SIMT
    if not fused_first:
      assert which is None
      call('immb_vgg_prologue', self.future_image, self.pred, self.pcs, B, R, 1, self.vgg_in.hi, self.vgg_in.lo, st)
    X = self.vgg_in
    prev_conv = None
    self._level_of = {nm: k for k, nm in enumerate(self.comp)}
    for kind, item, cin, size in self.vgg_seq:
      if kind == 'conv':
        _lib.TAG = 'fwd:vgg/%s' % item.name
        out = sel(item.out)
        if item.name == 'conv1_1' and fused_first:
          # Cin = 1: HBM-bound, exact-fp32 CUDA-core kernel straight from the RGB inputs (no patch tensor)
          call('immb_vgg_conv1_1_fused', self.future_image, None if which == 0 else self.pred, self.pcs, B, R,
               item.w, item.b, item.cout, item.out.hi, item.out.lo, 0 if which is None else which + 1, item.out.scale, st)
        else:
          self._conv_fwd(item, X, out.hi, out.lo, N=n, Y=out)
        X = out
        prev_conv = item.name
      else:
        O = sel(self.vgg_act[item])
        lvl = self._level_of.get(prev_conv) if (which is None and self.fuse_level_sums) else None
        if lvl is not None and cin % 4 == 0:
          # this level feeds a pool: its masked squared-difference sum rides in the pool kernel
          call('immb_maxpool2x2_fwd_levelsum', X.hi, X.lo, B, size, size, cin, O.hi, O.lo, self.mask, R,
               self.level_acc[lvl:], X.scale, O.scale, st)
          self._levels_done.add(lvl)
        else:
          call('immb_maxpool2x2_fwd', X.hi, X.lo, n, size, size, cin, O.hi, O.lo, X.scale, O.scale, st)
        X = O

  def _loss_fwd(self, training):
    """_colorization_reconstruction_loss (imm_model.py:111-151) + build_vgg16 (build_vgg16.py:14-35)."""
    if not self.vgg_loaded:
      raise _lib.ImmbError('VGG16 weights not loaded (load_vgg_caffe_dict)')
    st = _lib.stream_ptr()
    B, R = self.B, self.R
    self._levels_done = set()
    if self._gt_tower_forked:
      self._vgg_tower(1)                               # predicted half; the gt half was forked at the top of forward()
      self._join(self.gt_stream, 'gt_join')
      self._gt_tower_forked = False
    else:
      self._vgg_tower(None)
    for k, nm in enumerate(self.comp):
      if k in self._levels_done:
        continue
      if nm == 'input':
        call('immb_perceptual_level_sum', self.future_image, None, 3, self.pred, None, self.pcs, B, R, R, 3,
             self.mask, R, self.level_acc[k:], None, st)
      else:
        P = self.vgg_act[nm]
        h, w, C = P.hi.shape[1], P.hi.shape[2], P.hi.shape[3]
        g, p = P.half(0, B), P.half(1, B)
        call('immb_perceptual_level_sum', g.hi, g.lo, C, p.hi, p.lo, C, B, h, w, C, self.mask, R,
             self.level_acc[k:], P.scale, st)
    call('immb_perceptual_finalize', self.level_acc, self.level_counts, len(self.comp), self.agg,
         1 if training else 0, self.levels, self.rec_loss, self.coef, st)

  def loss_value(self):
    """total = reconstruction + sum_w 1e-5*0.5*||w||^2 (imm_model.py:395-400).  Device tensor [1]."""
    st = _lib.stream_ptr()
    self.sq.zero_()
    call('immb_adam_norms', self.flat_p, self.flat_g, self.n_flat, self.chunk_tensor, self.chunk_off,
         self.chunk_len, self.n_chunks, self.tensor_wd, 1.0, self.sq, self.sq[self.n_tensors:], st)
    call('immb_total_loss', self.rec_loss, self.sq[self.n_tensors:], self.tensor_wd, self.n_tensors,
         self.weights_loss, self.total_loss, self.h16_overflow if self.h16 else None, st)
    return self.total_loss

  # ------------------------------------------------------------------------------------------------
  # backward
  # ------------------------------------------------------------------------------------------------
  def _block_bwd(self, L, g, gcs, scratch=None):
    """g: gradient wrt the block output (after the optional x2 upsample), channel stride gcs.
    Returns the gradient wrt the block input [N,H,W,xcs] or None."""
    st = _lib.stream_ptr()
    _lib.TAG = 'bwd:%s/%s' % (L.prefix.split('/')[-2] if L.prefix.endswith('encoder') else L.prefix.split('/')[-1], L.name)
    npix = L.N * L.Ho * L.Wo
    if L.bn:
      if L.up2x:
        call('immb_upsample2x_bwd', g, L.N, L.Ho, L.Wo, L.cout, gcs, L.g_low, st)
        g, gcs = L.g_low, L.cout
      relu = 1 if L.relu else 0
      sc = self.bn_scratch if scratch is None else scratch
      if L.bwd_rows > 0:
        # sum(dz), sum(dz * xhat) were accumulated by the consumer's dgrad epilogue: fixed-order second level only
        call('immb_bn_stats_from_partials', sc, L.bwd_rows, L.cout, L.bsums, st)
        L.bwd_rows = 0
      else:
        call('immb_bn_bwd_reduce', g, gcs, L.y, L.cout, npix, L.cout, L.scale, L.shift, L.mean, L.invstd, relu,
             L.bsums, sc, sc.numel(), st)
      if L.db_partials is not None:
        # the per-block partial sums of the bias gradient stay in the layer's own buffer; ONE launch at the end of the
        # backward pass reduces them for all layers (no accumulator, no cast, no second-level launch per layer)
        call('immb_bn_bwd_apply', g, gcs, L.y, L.cout, npix, L.cout, L.scale, L.shift, L.mean, L.invstd, relu,
             L.bsums, L.dy.hi, L.dy.lo, L.dgamma, L.dbeta, L.dbias_acc, L.db_partials, L.db_partials.numel(), L.dy.scale,
             None, 1, st)
        direct_db = True
      else:
        direct_db = False
        call('immb_bn_bwd_apply', g, gcs, L.y, L.cout, npix, L.cout, L.scale, L.shift, L.mean, L.invstd, relu,
             L.bsums, L.dy.hi, L.dy.lo, L.dgamma, L.dbeta, L.dbias_acc, sc, sc.numel(), L.dy.scale, None, 0, st)
      dy = L.dy
    else:
      dy = g if isinstance(g, Planes) else None
      if dy is None:
        assert gcs == L.ycs
        call('immb_split_planes', g, L.dy.hi, L.dy.lo, npix * L.ycs, L.dy.scale, st)
        dy = L.dy
      call('immb_bias_grad', dy.hi, dy.lo, L.ycs, npix, L.cout, L.dbias_acc, dy.scale, st)
      direct_db = False
    if not direct_db:
      call('immb_cast_d2f', L.dbias_acc, L.db, L.cout, st)
    d = L.desc(x=L.x, y=dy)
    if self.wgrad_stream is not None:
      # dw is consumed by the optimiser only: the wgrad runs on its own stream behind the kernel that produced dy
      self._fork(self.wgrad_stream, ('wg', id(L)))
      with torch.cuda.stream(self.wgrad_stream):
        call('immb_conv2d_wgrad', d, L.x.hi, L.x.lo, dy.hi, dy.lo, L.dw, self.workspace, self.workspace.numel(),
             _lib.stream_ptr())
    else:
      call('immb_conv2d_wgrad', d, L.x.hi, L.x.lo, dy.hi, dy.lo, L.dw, self.workspace, self.workspace.numel(), st)
    if L.needs_dgrad:
      P = L.producer
      if P is not None and L.dgrad_stats_rows > 0 and self.fuse_bn_stats and self.training:
        sc = self.bn_scratch if scratch is None else scratch
        call('immb_conv2d_dgrad_bnreduce', d, dy.hi, dy.lo, L.wh.hi, L.wh.lo, L.dx, P.y, P.ycs, P.scale, P.shift,
             P.mean, P.invstd, 1 if P.relu else 0, sc, sc.numel(), st)
        P.bwd_rows = L.dgrad_stats_rows
      else:
        call('immb_conv2d_dgrad', d, dy.hi, dy.lo, L.w, L.wh.hi, L.wh.lo, L.dx, st)
      return L.dx
    return None

  def _loss_bwd(self):
    """Backward of the perceptual loss down to the renderer output; returns dy planes [B,R,R,n_out]."""
    st = _lib.stream_ptr()
    B, R = self.B, self.R
    level_of = {nm: k for k, nm in enumerate(self.comp)}
    g = None                       # gradient wrt the current activation (pred half), fp32 [B,h,w,C]
    fused_dy = None                # layer whose dy planes were already produced by the consumer's fused dgrad epilogue
    seq = self.vgg_seq
    for idx in range(len(seq) - 1, -1, -1):
      kind, item, cin, size = seq[idx]
      if kind == 'conv':
        L = item
        P = L.out
        coef = self.coef[level_of[L.name]:] if L.name in level_of else None
        if g is None and coef is None and fused_dy is not L:
          continue                  # above the deepest level used by the loss
        fg, fp = P.half(0, B), P.half(1, B)
        _lib.TAG = 'bwd:vgg/%s' % L.name
        if fused_dy is not L:
          call('immb_vgg_bwd_combine', g, fg.hi, fg.lo, fp.hi, fp.lo, B, size, size, L.cout, self.mask, R, coef,
               L.dy.hi, L.dy.lo, P.scale, L.dy.scale, st)
        fused_dy = None
        # producer of this conv's input: when it is a conv without a loss level, its dy = dgrad * [act > 0] is written
        # by this dgrad's epilogue (no fp32 gradient round trip, no combine launch)
        prev = seq[idx - 1] if idx > 0 else None
        d = L.desc(B, y=L.dy)
        if (L.name == 'conv1_1' and self.fuse_level_sums and self.engine != _lib.ENGINE_SIMT and R % 16 == 0
                and L.cout == 64 and 'input' in level_of):
          # conv1_1 dgrad + gray/normalise adjoint + the 'input' level's term: the renderer-output gradient in one kernel
          call('immb_vgg_conv1_1_bwd_fused', L.dy.hi, L.dy.lo, L.w, L.cout, self.future_image, self.pred, self.pcs,
               self.mask, self.coef[level_of['input']:], B, R, self.pred_dy.hi, self.pred_dy.lo, L.dy.scale,
               self.pred_dy.scale, st)
          return self.pred_dy
        if (prev is not None and prev[0] == 'conv' and prev[1].name not in level_of
                and _lib.lib().immb_conv2d_dgrad_relu_supported(d)):
          Lp = prev[1]
          call('immb_conv2d_dgrad_relu', d, L.dy.hi, L.dy.lo, L.wh.hi, L.wh.lo, Lp.out.half(1, B).hi, Lp.cout,
               Lp.dy.hi, Lp.dy.lo, Lp.dy.scale, st)
          fused_dy, g = Lp, None
        else:
          call('immb_conv2d_dgrad', d, L.dy.hi, L.dy.lo, L.w, L.wh.hi, L.wh.lo, L.dx, st)
          g = L.dx
      else:
        if g is None:
          continue
        # input of the pool = previous conv's activation (pred half)
        prev = seq[idx - 1][1]
        xin = prev.out.half(1, B)
        if self.fuse_level_sums and cin % 4 == 0:
          # pool backward + the producer's loss term / ReLU backward / operand split in one pass: dy of `prev` directly
          _lib.TAG = 'bwd:vgg/%s' % prev.name
          fgp = prev.out.half(0, B)
          coef_p = self.coef[level_of[prev.name]:] if prev.name in level_of else None
          call('immb_maxpool2x2_bwd_combine', g, fgp.hi, fgp.lo, xin.hi, xin.lo, B, size, size, cin, self.mask, R,
               coef_p, prev.dy.hi, prev.dy.lo, xin.scale, prev.dy.scale, st)
          fused_dy, g = prev, None
        else:
          call('immb_maxpool2x2_bwd', g, xin.hi, xin.lo, B, size, size, cin, self.g_pool[item], st)
          g = self.g_pool[item]
    coef_in = self.coef[level_of['input']:] if 'input' in level_of else None
    if coef_in is None:
      raise _lib.ImmbError("perceptual.comp without 'input' is not built")
    call('immb_pred_grad', self.future_image, self.pred, self.pcs, self.mask, coef_in, g, 1, B, R,
         self.pred_dy.hi, self.pred_dy.lo, st)
    return self.pred_dy

  def backward(self, allreduce=None):
    """Gradients of (reconstruction loss) wrt every trainable tensor; the L2 term is added in the optimiser.
    allreduce (N > 1): callable(flat tensor) = sum over replicas (cnn_train_multi.py:66-106).  The flat gradient buffer is
    reduced in two buckets in backward order: the renderer's gradients (the tail of the buffer, complete after the
    renderer's backward) go out on a communication stream while the two encoders are still running their backward; the
    encoders' bucket follows the last weight-gradient kernel.  The optimiser waits for both."""
    st = _lib.stream_ptr()
    B, K = self.B, self.K
    self.bwd_pool.zero_()
    g = self._loss_bwd()
    gcs = self.pcs
    for L in reversed(self.ren_layers):
      g = self._block_bwd(L, g, gcs)
      gcs = L.xcs
    comm = None
    early_ren_db = allreduce is not None and self.overlap_allreduce and self._db_table is not None
    if early_ren_db:
      # the renderer's bias gradients belong to the first bucket: reduce their partials now (items are in layer order,
      # the renderer's are the tail of the table)
      call('immb_reduce_partials_multi', self._db_table_ren, self._db_n_ren, self._db_blocks_ren, _lib.stream_ptr())
    if allreduce is not None and self.overlap_allreduce:
      if self.comm_stream is None:
        self.comm_stream = torch.cuda.Stream(device=self.dev)
      comm = self.comm_stream
      self._fork(comm, 'ar_fork_main')                       # renderer db / dgamma / dbeta (main stream)
      if self.wgrad_stream is not None:
        ev = self._event('ar_fork_wg')
        ev.record(self.wgrad_stream)                         # renderer dw (weight-gradient stream)
        comm.wait_event(ev)
      with torch.cuda.stream(comm):
        allreduce(self.flat_g[self.ren_grad_offset:])
    dJ = g                                             # [B,16,16,Cj]
    # pose branch: Gaussian maps -> mu -> softmax marginals -> heatmaps (imm_model.py:252-274)
    S = self.enc_out_size
    def pose_branch_bwd(scratch):
      call('immb_softargmax_gauss_bwd', dJ, self.Cj, self.enc_feat, self.mu, self.py, self.px, B, S, K, 16,
           self.inv_std, self.g_heat, self.Kp, _lib.stream_ptr())
      gp = self._block_bwd(self.pose_conv, self.g_heat, self.Kp, scratch)
      gcs_p = self.enc_feat
      for L in reversed(self.enc_layers['pose_encoder']):
        gp = self._block_bwd(L, gp, gcs_p, scratch)
        gcs_p = L.xcs

    if self.pose_stream is not None:
      self._fork(self.pose_stream, 'bwd_fork')
      with torch.cuda.stream(self.pose_stream):
        pose_branch_bwd(self.bn_scratch_pose)
    else:
      pose_branch_bwd(None)
    # image branch
    gi, gcs_i = dJ, self.Cj
    if self.enc_out_size != 16:
      call('immb_resize_ac_bwd', dJ, self.Cj, B, S, S, self.enc_feat, 16, 16, self.g_enc_resized, st)
      gi, gcs_i = self.g_enc_resized, self.enc_feat
    for L in reversed(self.enc_layers['image_encoder']):
      gi = self._block_bwd(L, gi, gcs_i)
      gcs_i = L.xcs
    if self.pose_stream is not None:
      self._join(self.pose_stream, 'bwd_join')
    if self.wgrad_stream is not None:
      self._join(self.wgrad_stream, 'wg_join')
    if self._db_table is not None:
      if early_ren_db:
        call('immb_reduce_partials_multi', self._db_table_enc, self._db_n_enc, self._db_blocks_enc, _lib.stream_ptr())
      else:
        call('immb_reduce_partials_multi', self._db_table, self._db_n, self._db_blocks, _lib.stream_ptr())
    if allreduce is not None:
      if comm is not None:
        allreduce(self.flat_g[:self.ren_grad_offset])
        self._join(comm, 'ar_join')
      else:
        allreduce(self.flat_g)

  # ------------------------------------------------------------------------------------------------
  # optimiser  (cnn_train_multi.py:93-98,232-241; scripts/train.py:92-98)
  # ------------------------------------------------------------------------------------------------
  def learning_rate(self, start_val=1e-3, step=100000, decay=0.95, lr_multiple=1.0):
    return lr_multiple * start_val * decay ** math.floor(self.global_step / float(step))

  def _optimizer_kernels(self, clip_value, lr_t, beta1, beta2, eps, lr_t_dev=None):
    """+wd*w -> per-tensor clip_by_norm -> TF Adam -> repack the tensor-core weight planes; total loss value."""
    st = _lib.stream_ptr()
    gscale = 1.0 / float(self.world_size)
    clip = float(clip_value) if clip_value is not None else 0.0
    self.sq.zero_()
    nt = self.n_tensors
    call('immb_adam_norms', self.flat_p, self.flat_g, self.n_flat, self.chunk_tensor, self.chunk_off,
         self.chunk_len, self.n_chunks, self.tensor_wd, gscale, self.sq, self.sq[nt:], st)
    call('immb_total_loss', self.rec_loss, self.sq[nt:], self.tensor_wd, nt, self.weights_loss, self.total_loss,
         self.h16_overflow if self.h16 else None, st)
    amax = None
    if self.h16:
      self.w_amax.zero_()
      amax = self.w_amax                 # max |p| of the updated tensors: the scale of their fp16 weight planes
    if lr_t_dev is None:
      call('immb_adam_apply', self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.n_flat, self.chunk_tensor,
           self.chunk_off, self.chunk_len, self.n_chunks, self.tensor_wd, gscale, self.sq, clip, lr_t, beta1, beta2,
           eps, amax, st)
    else:
      call('immb_adam_apply_dev', self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.n_flat, self.chunk_tensor,
           self.chunk_off, self.chunk_len, self.n_chunks, self.tensor_wd, gscale, self.sq, clip, lr_t_dev, beta1,
           beta2, eps, amax, st)
    self.repack_weights(amax_current=True)

  def _next_lr_t(self, lr, beta1, beta2):
    if lr is None:
      lr = self.learning_rate()
    self.adam_t += 1
    self.last_lr = lr
    self.adam_betas = (float(beta1), float(beta2))      # checkpointed as beta1_power / beta2_power (TF AdamOptimizer)
    return lr * math.sqrt(1.0 - beta2 ** self.adam_t) / (1.0 - beta1 ** self.adam_t)

  def optimizer_step(self, clip_value=1.0, lr=None, beta1=0.9, beta2=0.999, eps=1e-8, allreduce=None):
    """mean over replicas (all-reduce of the flat gradient buffer, unless backward(allreduce=...) already did it) ->
    +wd*w -> per-tensor clip_by_norm -> TF Adam -> repack the tensor-core weight planes.  Also produces the total loss."""
    if allreduce is not None:
      allreduce(self.flat_g)
    lr_t = self._next_lr_t(lr, beta1, beta2)
    self._optimizer_kernels(clip_value, lr_t, beta1, beta2, eps)
    self.global_step += 1.0
    return self.last_lr

  def train_step(self, image, future_image, mask=None, clip_value=1.0, lr_multiple=1.0, allreduce=None, lr=None,
                 beta1=0.9, beta2=0.999, eps=1e-8):
    """One iteration of train_loop's hot loop (cnn_train_multi.py:445-460): fwd + loss + bwd + update.
    With use_graph the whole step (3 streams, ~370 launches) is captured once into CUDA graphs and replayed: inputs are
    copied into static buffers, the step-dependent Adam scalar is refreshed in device memory, and (N > 1) the NCCL
    all-reduce runs between the forward+backward graph and the optimiser graph."""
    if lr is None:
      lr = self.learning_rate(lr_multiple=lr_multiple)
    if self.h16:
      # exponents of the training mode (calibrated on first use; restored after an interleaved evaluation pass) -- done
      # here, eagerly, so that neither graph capture nor graph replay ever contains a mode switch
      self._enter_scale_mode(image, future_image, mask, True, True)
    if self.use_graph:
      key = (None if clip_value is None else float(clip_value), beta1, beta2, eps, allreduce is not None,
             mask is not None)
      if self._graphs is not None and self._graph_key != key:
        self._graphs = None                                     # hyper-parameters baked into the graph changed
      if self._graphs is None and self._graph_warm >= 2:
        try:
          self._allreduce_fn = allreduce
          self._capture_graphs(key, image, future_image, mask)
        except Exception as e:      # e.g. a foreign thread touching CUDA during capture: keep training, eagerly
          import warnings
          warnings.warn('CUDA-graph capture of the training step failed (%s); continuing with eager launches' % (e,))
          self.use_graph, self._graphs = False, None
      if self._graphs is not None:
        self.g_image.copy_(image, non_blocking=True)
        self.g_future.copy_(future_image, non_blocking=True)
        if mask is not None:
          self.g_mask.copy_(mask, non_blocking=True)
        self.d_lr_t.fill_(self._next_lr_t(lr, beta1, beta2))    # by-value upload: no host buffer to race with
        g_fb, g_opt = self._graphs
        g_fb.replay()
        if g_opt is not None:                                   # NCCL could not be captured: all-reduce between two graphs
          allreduce(self.flat_g)
          g_opt.replay()
        self.global_step += 1.0
        self.graph_replays += 1
        return self.total_loss
      self._graph_warm += 1          # eager warm-up steps: kernel attributes configured, events created
    self.forward(image, future_image, mask, training=True, build_loss=True)
    self.backward(allreduce=allreduce)
    self.optimizer_step(clip_value, lr=lr, beta1=beta1, beta2=beta2, eps=eps)
    return self.total_loss

  def release_graphs(self):
    """Drop the captured step graphs (the next train_step re-captures).  Required before the NCCL process group is
    destroyed when the all-reduce was captured into the graph: a live graph holding NCCL kernels keeps the communicator
    busy and destroy_process_group() never returns."""
    if getattr(self, '_graphs', None) is not None:
      torch.cuda.synchronize(self.dev)
      for g in self._graphs:
        if g is not None:
          g.reset()
      self._graphs = None
      self._graph_key = None

  def _capture_graphs(self, key, image, future_image, mask):
    clip_value, beta1, beta2, eps, has_allreduce, has_mask = key
    dev = self.dev
    self.g_image = torch.empty_like(image, device=dev)
    self.g_future = torch.empty_like(future_image, device=dev)
    self.g_mask = torch.empty_like(mask, device=dev) if has_mask else None
    self.g_image.copy_(image)
    self.g_future.copy_(future_image)
    if has_mask:
      self.g_mask.copy_(mask)
    self.d_lr_t = torch.zeros(1, dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    saved = (self.flat_p.clone(), self.flat_m.clone(), self.flat_v.clone(), self.flat_bn.clone(), self.agg.clone())
    cap = torch.cuda.Stream(device=dev)
    g_fb, g_opt = torch.cuda.CUDAGraph(), None
    n0 = _lib.launch_count()
    one_graph = has_allreduce and self.graph_nccl
    _dbg('capture begin (one_graph=%s, overlap=%s)' % (one_graph, self.overlap_allreduce))
    if one_graph:
      # the whole step incl. the two bucketed NCCL all-reduces as ONE graph (NCCL >= 2.9 collectives are capturable)
      try:
        with torch.cuda.graph(g_fb, stream=cap, capture_error_mode='thread_local'):
          self.forward(self.g_image, self.g_future, self.g_mask, training=True, build_loss=True)
          self.backward(allreduce=self._allreduce_fn)
          self._optimizer_kernels(clip_value, 0.0, beta1, beta2, eps, lr_t_dev=self.d_lr_t)
      except Exception as e:
        import warnings
        warnings.warn('capturing the NCCL all-reduce into the step graph failed (%s); using two graphs' % (e,))
        one_graph, self.graph_nccl = False, False
        g_fb = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        torch.cuda.synchronize(dev)
    if not one_graph:
      with torch.cuda.graph(g_fb, stream=cap, capture_error_mode='thread_local'):
        self.forward(self.g_image, self.g_future, self.g_mask, training=True, build_loss=True)
        self.backward()
        if not has_allreduce:
          self._optimizer_kernels(clip_value, 0.0, beta1, beta2, eps, lr_t_dev=self.d_lr_t)
      if has_allreduce:
        g_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_opt, stream=cap, capture_error_mode='thread_local'):
          self._optimizer_kernels(clip_value, 0.0, beta1, beta2, eps, lr_t_dev=self.d_lr_t)
    torch.cuda.synchronize(dev)
    # capture does not execute anything, but be explicit that model state is exactly what it was
    for dst, src in zip((self.flat_p, self.flat_m, self.flat_v, self.flat_bn, self.agg), saved):
      dst.copy_(src)
    self._graphs, self._graph_key = (g_fb, g_opt), key
    _dbg('capture done (%s)' % ('one graph' if g_opt is None else 'two graphs + all-reduce between'))
    self.graph_launches_per_step = _lib.launch_count() - n0     # kernels recorded into the graphs (= launched per replay)

  # ------------------------------------------------------------------------------------------------
  # introspection
  # ------------------------------------------------------------------------------------------------
  def dtype_string(self):
    """The arithmetic the conv engine computes in (bench.py's `dtype`)."""
    if self.h16:
      n16 = sum(1 for L in list(self.layers.values()) + self._vgg_convs() if L.h16)
      return ('f16x3 (scaled fp16 split operands hi + lo*2^-11 with one power-of-two scale per tensor; tensor-core products '
              'hi*hi + (hi*lo + lo*hi)*2^-11 at kind::f16, fp32 accumulate: 22 significant bits per operand; %d of %d convs -- '
              'every 3x3 layer (stride 1 and 2) and the frozen VGG16 tower (2 passes: weights rounded to fp16 planes) -- the '
              '7x7 first layers and the 1x1 heat-map conv run error-compensated 3xTF32)' % (n16, len(self.layers) + len(self._vgg_convs())))
    if self.precision == _lib.PREC_TF32:
      return 'tf32 (single pass; does not meet the parity bar)'
    return ('tf32x3 (fp32 storage; error-compensated TF32 tensor-core products hi*hi+hi*lo+lo*hi, fp32 accumulate; '
            'the frozen VGG16 tower uses weights rounded to TF32 at load and 2 passes)')

  def engine_table(self):
    out = OrderedDict()
    for key, L in self.layers.items():
      out[key] = L.engines()
    for kind, item, cin, size in self.vgg_seq:
      if kind == 'conv':
        out['vgg16/' + item.name] = item.engines()
    return out

  def gaussian_maps(self, mu, size):
    """get_gaussian_maps(mu, [size,size], inv_std, 'rot') (imm_model.py:34-78) -> [B,size,size,K]."""
    B, K = mu.shape[0], mu.shape[1]
    out = torch.empty((B, size, size, K), dtype=torch.float32, device=self.dev)
    call('immb_gaussian_maps', mu.contiguous(), B, K, size, self.inv_std, out, _lib.stream_ptr())
    return out


# ---- device scoping ---------------------------------------------------------------------------------------
# Every public entry point runs with the engine's device current, so an engine on 'cuda:1' enqueues on cuda:1's streams
# even when the caller never called torch.cuda.set_device (the C ABI launches on the CURRENT device's stream).
def _on_engine_device(fn):
  import functools

  @functools.wraps(fn)
  def wrapped(self, *args, **kwargs):
    with torch.cuda.device(self.dev):
      return fn(self, *args, **kwargs)
  return wrapped


def _init_on_device(fn):
  import functools

  @functools.wraps(fn)
  def wrapped(self, config, batch, image_size=128, device='cuda:0', *args, **kwargs):
    if not torch.cuda.is_available():
      return fn(self, config, batch, image_size, device, *args, **kwargs)       # raises: there is no CPU fallback
    with torch.cuda.device(torch.device(device)):
      return fn(self, config, batch, image_size, device, *args, **kwargs)
  return wrapped


IMMEngine.__init__ = _init_on_device(IMMEngine.__init__)
for _name in ('init_parameters', 'load_state', 'load_vgg_caffe_dict', 'load_vgg_hwio', 'repack_weights', 'forward', 'backward',
              'optimizer_step', 'train_step', 'release_graphs', 'run_image_encoder', 'run_pose_branch', 'run_renderer',
              'load_joint', 'loss_value', 'gaussian_maps'):
  setattr(IMMEngine, _name, _on_engine_device(getattr(IMMEngine, _name)))
del _name
