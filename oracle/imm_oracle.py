"""
ORACLE -- TEST INFRASTRUCTURE ONLY.  NOT PRODUCT CODE.

CPU restatement (PyTorch-CPU, fp32 or fp64, autograd for the backward pass) of the
arithmetic of the IMM training hot path of tomasjakab/imm.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module.  Nothing under `imm_b200/` imports it.

PARITY UNPINNED: the reference has no tests, golden vectors or fixtures, and its
arithmetic lives in `tensorflow-gpu==1.10.0` (requirements.txt:1), which is not
under /root/reference and cannot be installed here (Python 3.12, no wheel, no
network).  This file therefore restates TF-1.10's published op semantics at the
reference's call sites; every TF-semantic it depends on ([TF-sem]) is pinned by a
hand-derived known-answer test in tests/test_oracle_kat.py.

All citations are file:line under /root/reference.

Layout: tensors are NHWC at every function boundary (as in the reference);
conv weights are HWIO (base_model.py:110).  Parameters live in a flat dict keyed
by the reference's TF variable names ("Checkpoint layout").
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# hyper-parameters fixed by the reference
# ----------------------------------------------------------------------------
BN_EPS = 1e-3          # tf.layers.batch_normalization default epsilon  [TF-sem]
BN_MOMENTUM = 0.99     # tf.layers.batch_normalization default momentum [TF-sem]
WD = 1e-5              # base_model.py:65
INIT_STD = 0.01        # base_model.py:66
PERCEPTUAL_WS = [100.0, 1.6, 2.3, 1.8, 2.8, 100.0]   # imm_model.py:131
VGG_MEAN = 114.451     # build_vgg16.py:26
VGG_LAYERS = [('conv1_1', 64), ('conv1_2', 64), 'pool1',
              ('conv2_1', 128), ('conv2_2', 128), 'pool2',
              ('conv3_1', 256), ('conv3_2', 256), ('conv3_3', 256), 'pool3',
              ('conv4_1', 512), ('conv4_2', 512), ('conv4_3', 512), 'pool4',
              ('conv5_1', 512), ('conv5_2', 512), ('conv5_3', 512), 'pool5']   # vgg16.py:338-373


# ----------------------------------------------------------------------------
# TF-1.10 op restatements
# ----------------------------------------------------------------------------
def same_pad(in_size, k, stride):
  """[TF-sem] SAME padding: out = ceil(in/stride); pad_total = max((out-1)*stride + k - in, 0);
  pad_before = pad_total // 2, pad_after = pad_total - pad_before (extra goes AFTER)."""
  out = -(-in_size // stride)
  total = max((out - 1) * stride + k - in_size, 0)
  return total // 2, total - total // 2


# Optional operand filter (test-only): lets tests emulate tensor-core operand rounding
# (e.g. TF32 round-to-nearest) inside the oracle to bound the expected GPU-vs-oracle gap.
OPERAND_FILTER = None


def round_tf32(t):
  """Round-to-nearest-even of fp32 values to TF32 (10 explicit mantissa bits), straight-through grad."""
  if t.dtype != torch.float32:
    return t
  with torch.no_grad():
    i = t.detach().contiguous().view(torch.int32)
    r = (i + 0x0FFF + ((i >> 13) & 1)) & ~0x1FFF
    q = r.view(torch.float32)
  return t + (q - t).detach()


def conv2d_same(x, w, b=None, stride=1):
  """tf.nn.conv2d(x, w, strides=[1,s,s,1], padding='SAME') (+ tf.nn.bias_add).
  nn_utils.py:100,108.  x NHWC, w HWIO, cross-correlation."""
  if OPERAND_FILTER is not None:
    x, w = OPERAND_FILTER(x), OPERAND_FILTER(w)
  kh, kw = w.shape[0], w.shape[1]
  pt, pb = same_pad(x.shape[1], kh, stride)
  pl, pr = same_pad(x.shape[2], kw, stride)
  xn = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
  y = F.conv2d(xn, w.permute(3, 2, 0, 1), bias=b, stride=stride)
  return y.permute(0, 2, 3, 1)


def batch_norm(x, gamma, beta, moving_mean, moving_var, training):
  """tf.layers.batch_normalization(x, training=..., fused=True)  (nn_utils.py:201-202).
  [TF-sem] axis=-1, eps=1e-3, momentum=0.99.  Training: normalise with the BIASED batch
  variance over (B,H,W); the moving variance is updated with the BESSEL-corrected one
  (fused kernel output, `_bessels_correction_test_only=True` keeps it);
  update: mv -= (mv - v) * (1 - momentum).  Returns (y, new_moving_mean, new_moving_var)."""
  if training:
    n = x.shape[0] * x.shape[1] * x.shape[2]
    mean = x.mean(dim=(0, 1, 2))
    var = ((x - mean) ** 2).mean(dim=(0, 1, 2))
    y = (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta
    with torch.no_grad():
      var_unb = var * (n / max(n - 1.0, 1.0))
      new_mm = moving_mean - (moving_mean - mean) * (1.0 - BN_MOMENTUM)
      new_mv = moving_var - (moving_var - var_unb) * (1.0 - BN_MOMENTUM)
    return y, new_mm, new_mv
  y = (x - moving_mean) * torch.rsqrt(moving_var + BN_EPS) * gamma + beta
  return y, moving_mean, moving_var


def resize_bilinear(x, out_hw, align_corners=False):
  """tf.image.resize_images / tf.image.resize_bilinear, TF-1.x kernel.
  [TF-sem] align_corners=False: scale = in/out, src = dst*scale (LEGACY: no half-pixel
  offset); align_corners=True: scale = (in-1)/(out-1).  lower = floor(src),
  upper = min(lower+1, in-1), linear interpolation.  imm_model.py:175,334,409."""
  def axis_weights(n_in, n_out):
    if align_corners and n_out > 1:
      scale = (n_in - 1) / float(n_out - 1)
    else:
      scale = n_in / float(n_out)
    src = torch.arange(n_out, dtype=torch.float64) * scale
    lo = torch.floor(src).long().clamp(max=n_in - 1)
    hi = (lo + 1).clamp(max=n_in - 1)
    frac = (src - lo.double()).to(x.dtype)
    return lo, hi, frac
  H, W = x.shape[1], x.shape[2]
  lo_h, hi_h, fh = axis_weights(H, out_hw[0])
  lo_w, hi_w, fw = axis_weights(W, out_hw[1])
  top = x[:, lo_h]
  bot = x[:, hi_h]
  def lerp_w(t):
    l = t[:, :, lo_w]
    r = t[:, :, hi_w]
    return l + (r - l) * fw.view(1, 1, -1, 1)
  t = lerp_w(top)
  b_ = lerp_w(bot)
  return t + (b_ - t) * fh.view(1, -1, 1, 1)


def max_pool_2x2(x):
  """tf.nn.max_pool(ksize 2, stride 2, SAME) on even sizes (selfsup/ops.py:20)."""
  y = F.max_pool2d(x.permute(0, 3, 1, 2), kernel_size=2, stride=2, ceil_mode=True)
  return y.permute(0, 2, 3, 1)


def clip_by_norm(g, clip):
  """[TF-sem] tf.clip_by_norm: g * clip / max(||g||_2, clip).  cnn_train_multi.py:98,237."""
  l2 = torch.sqrt((g * g).sum())
  return g * clip / torch.maximum(l2, torch.tensor(clip, dtype=g.dtype))


def adam_step(var, g, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
  """[TF-sem] tf.train.AdamOptimizer (scripts/train.py:98): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
  m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g; var -= lr_t*m/(sqrt(v)+eps)  (eps OUTSIDE the
  bias correction, unlike torch.optim.Adam).  t counts from 1."""
  lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
  m = beta1 * m + (1.0 - beta1) * g
  v = beta2 * v + (1.0 - beta2) * g * g
  var = var - lr_t * m / (torch.sqrt(v) + eps)
  return var, m, v


def learning_rate(global_step, start_val=1e-3, step=100000, decay=0.95, lr_multiple=1.0):
  """scripts/train.py:92-96: lr_multiple * exponential_decay(start, gs, step, decay, staircase=True).
  NB the reference initialises global_step to --reset-global-step, default -1 (train.py:87-89,192),
  so the very first step runs with floor(-1/1e5) = -1 -> lr = 1e-3/0.95."""
  return lr_multiple * start_val * decay ** math.floor(global_step / float(step))


# ----------------------------------------------------------------------------
# model pieces
# ----------------------------------------------------------------------------
def get_gaussian_maps(mu, shape_hw, inv_std, mode='rot'):
  """imm_model.py:34-78.  mu [B,K,2] (y,x) in [-1,1] -> [B,H,W,K]."""
  mu_y, mu_x = mu[:, :, 0:1], mu[:, :, 1:2]
  y = torch.linspace(-1.0, 1.0, shape_hw[0], dtype=mu.dtype)
  x = torch.linspace(-1.0, 1.0, shape_hw[1], dtype=mu.dtype)
  if mode in ('rot', 'flat'):
    mu_y, mu_x = mu_y.unsqueeze(-1), mu_x.unsqueeze(-1)
    y = y.view(1, 1, -1, 1)
    x = x.view(1, 1, 1, -1)
    dist = ((y - mu_y) ** 2 + (x - mu_x) ** 2) * inv_std ** 2
    g_yx = torch.exp(-dist) if mode == 'rot' else torch.exp(-torch.pow(dist + 1e-5, 0.25))
  elif mode == 'ankush':
    y = y.view(1, 1, -1)
    x = x.view(1, 1, -1)
    g_y = torch.exp(-torch.sqrt(1e-4 + torch.abs((mu_y - y) * inv_std)))
    g_x = torch.exp(-torch.sqrt(1e-4 + torch.abs((mu_x - x) * inv_std)))
    g_yx = g_y.unsqueeze(3) @ g_x.unsqueeze(2)
  else:
    raise ValueError('Unknown mode: ' + str(mode))
  return g_yx.permute(0, 2, 3, 1)


def get_coord(x, other_axis, axis_size):
  """imm_model.py:252-259: mean over the other axis -> 1-D softmax -> expectation."""
  g_c_prob = x.mean(dim=other_axis)                  # B,S,K
  g_c_prob = torch.softmax(g_c_prob, dim=1)
  coord_pt = torch.linspace(-1.0, 1.0, axis_size, dtype=x.dtype).view(1, axis_size, 1)
  g_c = (g_c_prob * coord_pt).sum(dim=1)
  return g_c, g_c_prob


class State(object):
  """Parameters + non-trainable state keyed by TF variable name, plus the model config."""

  def __init__(self, n_maps=10, image_size=128, dtype=torch.float32, n_filters=32,
               n_filters_render=32, gauss_std=0.1, gauss_mode='rot', renderer_stride=2,
               min_res=16, perceptual_comp=('input', 'conv1_2', 'conv2_2', 'conv3_2', 'conv4_2', 'conv5_2'),
               channels_bug_fix=True, loss_mask=True):
    self.n_maps = n_maps
    self.image_size = image_size
    self.dtype = dtype
    self.n_filters = n_filters
    self.n_filters_render = n_filters_render
    self.gauss_std = gauss_std
    self.gauss_mode = gauss_mode
    self.renderer_stride = renderer_stride
    self.min_res = min_res
    self.perceptual_comp = list(perceptual_comp)
    self.channels_bug_fix = channels_bug_fix
    self.loss_mask = loss_mask
    self.params = OrderedDict()       # trainable: name -> tensor (requires_grad set by caller)
    self.buffers = OrderedDict()      # BN moving stats, *_agg
    self.vgg = OrderedDict()          # SelfSupReconstructionLoss/vgg16/<name>/{weights,biases}
    self.adam_m = OrderedDict()
    self.adam_v = OrderedDict()
    self.adam_t = 0
    self.global_step = -1.0           # scripts/train.py:87-89,192 (see learning_rate())

  def clone(self, dtype=None):
    dtype = dtype or self.dtype
    s = State.__new__(State)
    s.__dict__.update({k: v for k, v in self.__dict__.items()
                       if k not in ('params', 'buffers', 'vgg', 'adam_m', 'adam_v')})
    s.dtype = dtype
    for name in ('params', 'buffers', 'vgg', 'adam_m', 'adam_v'):
      setattr(s, name, OrderedDict((k, v.detach().clone().to(dtype)) for k, v in getattr(self, name).items()))
    return s


def encoder_spec(n_filters):
  """imm_model.py:182-217 -> [(name, k, stride, cout)]."""
  f = n_filters
  return [('conv_1', 7, 1, f), ('conv_2', 3, 1, f),
          ('conv_3', 3, 2, 2 * f), ('conv_4', 3, 1, 2 * f),
          ('conv_5', 3, 2, 4 * f), ('conv_6', 3, 1, 4 * f),
          ('conv_7', 3, 2, 8 * f), ('conv_8', 3, 1, 8 * f)]


def renderer_spec(n_filters_render, final_res, n_final_out, start_res=16):
  """imm_model.py:154-179 -> [(name, cout, batch_norm, relu, upsample_after)]."""
  filters = n_filters_render * 8
  size = start_res
  conv_id = 1
  spec = []
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


def trunc_normal(gen, shape, std):
  """tf.truncated_normal_initializer: N(0,std) re-drawn outside 2 std (nn_utils.py:47)."""
  t = torch.empty(shape, dtype=torch.float32)
  torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2 * std, b=2 * std, generator=gen)
  return t


def init_state(state, seed=0, vgg_data=None):
  """Creates every variable of SURVEY 8a 'Checkpoint layout' with the reference initialisers:
  w trunc-normal(0.01), b 0 (nn_utils.py:47,106); gamma 1, beta 0, moving_mean 0, moving_var 1
  (tf.layers defaults); *_agg = ws (imm_model.py:131,144)."""
  gen = torch.Generator().manual_seed(seed)
  P, Bf = state.params, state.buffers

  def add_conv(scope, name, k, cin, cout, bn):
    P['%s/%s/%s/w' % (scope, name, name)] = trunc_normal(gen, (k, k, cin, cout), INIT_STD)
    P['%s/%s/%s/b' % (scope, name, name)] = torch.zeros(cout)
    if bn:
      P['%s/%s/batch_normalization/gamma' % (scope, name)] = torch.ones(cout)
      P['%s/%s/batch_normalization/beta' % (scope, name)] = torch.zeros(cout)
      Bf['%s/%s/batch_normalization/moving_mean' % (scope, name)] = torch.zeros(cout)
      Bf['%s/%s/batch_normalization/moving_variance' % (scope, name)] = torch.ones(cout)

  for enc in ('model/image_encoder/encoder', 'model/pose_encoder/encoder'):
    cin = 3
    for name, k, _, cout in encoder_spec(state.n_filters):
      add_conv(enc, name, k, cin, cout, True)
      cin = cout
  add_conv('model/pose_encoder', 'conv_1', 1, 8 * state.n_filters, state.n_maps, False)
  n_out = 3 + (len(state.perceptual_comp) if state.channels_bug_fix else 0)
  cin = 8 * state.n_filters + state.n_maps
  for name, cout, bn, _, _ in renderer_spec(state.n_filters_render, state.image_size, n_out):
    add_conv('model/renderer', name, 3, cin, cout, bn)
    cin = cout
  for k, nm in enumerate(state.perceptual_comp):
    Bf['SelfSupReconstructionLoss/%s_agg' % nm] = torch.tensor(PERCEPTUAL_WS[k])
  if vgg_data is None:
    vgg_data = synthetic_vgg_caffe_dict(seed + 1)
  state.vgg = load_vgg_params(vgg_data)
  for d in (state.params, state.buffers, state.vgg):
    for k_ in list(d.keys()):
      d[k_] = d[k_].to(state.dtype)
  for k_, v in state.params.items():
    state.adam_m[k_] = torch.zeros_like(v)
    state.adam_v[k_] = torch.zeros_like(v)
  return state


def synthetic_vgg_caffe_dict(seed=1):
  """The real vgg16.caffemodel.h5 (configs/paths/default.yaml:6) is a download and is not
  available offline.  This generates seeded weights in the SAME dict layout deepdish returns
  (build_vgg16.py:16; vgg16.py:19-40,76-87): data[name]['0'] = W [O,I,kh,kw] (Caffe OIHW),
  data[name]['1'] = bias [O]; data['batch_'+name]['0'|'1'|'2'] = mean*s, var*s, s."""
  rng = np.random.RandomState(seed)
  data = {}
  cin = 1
  for item in VGG_LAYERS:
    if isinstance(item, str):
      continue
    name, cout = item
    std = math.sqrt(2.0 / (9 * cin))
    data[name] = {'0': (rng.randn(cout, cin, 3, 3) * std).astype(np.float32),
                  '1': (rng.randn(cout) * 0.05).astype(np.float32)}
    s = np.float32(2.0)
    data['batch_' + name] = {'0': (rng.randn(cout) * 0.1).astype(np.float32) * s,
                             '1': (0.5 + rng.rand(cout)).astype(np.float32) * s,
                             '2': np.array([s], dtype=np.float32)}
    cin = cout
  return data


def load_vgg_params(data):
  """vgg16.py:17-47 (weights) and :74-92 (biases) with pre_adjust_batch_norm=True
  (build_vgg16.py:30): OIHW->HWIO transpose (:27); BGR flip only when Cin==3 (:28-29);
  W /= sigma, b = (b - mu)/sigma with sigma = sqrt(1e-5 + bn['1']/bn['2']), mu = bn['0']/bn['2']."""
  out = OrderedDict()
  for item in VGG_LAYERS:
    if isinstance(item, str):
      continue
    name, _ = item
    W = np.array(data[name]['0'], dtype=np.float32).copy().transpose(2, 3, 1, 0)
    if name == 'conv1_1' and W.shape[2] == 3:
      W = W[:, :, ::-1]
    bias = np.array(data[name]['1'], dtype=np.float32).copy()
    bn_name = 'batch_' + name
    if bn_name in data:
      bn = data[bn_name]
      sigma = np.sqrt(1e-5 + bn['1'] / bn['2'])
      mu = bn['0'] / bn['2']
      W = W / sigma
      bias = (bias - mu) / sigma
    out['SelfSupReconstructionLoss/vgg16/%s/weights' % name] = torch.from_numpy(np.ascontiguousarray(W, dtype=np.float32))
    out['SelfSupReconstructionLoss/vgg16/%s/biases' % name] = torch.from_numpy(np.ascontiguousarray(bias, dtype=np.float32))
  return out


def conv_block(state, new_buffers, scope, name, x, stride, bn, relu, training):
  """IMMModel.conv (imm_model.py:103-108) -> BaseModel.conv_block (base_model.py:96-117) ->
  nnu.conv_block (nn_utils.py:151-210): conv -> bias -> [BN] -> [ReLU]."""
  P = state.params
  w = P['%s/%s/%s/w' % (scope, name, name)]
  b = P['%s/%s/%s/b' % (scope, name, name)]
  y = conv2d_same(x, w, b, stride)
  if bn:
    pre = '%s/%s/batch_normalization/' % (scope, name)
    y, mm, mv = batch_norm(y, P[pre + 'gamma'], P[pre + 'beta'],
                           state.buffers[pre + 'moving_mean'], state.buffers[pre + 'moving_variance'], training)
    new_buffers[pre + 'moving_mean'] = mm
    new_buffers[pre + 'moving_variance'] = mv
  if relu:
    y = torch.relu(y)
  return y


def encoder(state, new_buffers, scope, x, training):
  """imm_model.py:182-217.  Returns the 4 block outputs."""
  feats = []
  for i, (name, _, stride, _) in enumerate(encoder_spec(state.n_filters)):
    x = conv_block(state, new_buffers, scope, name, x, stride, True, True, training)
    if i % 2 == 1:
      feats.append(x)
  return feats


def render_sizes(state, max_size):
  """imm_model.py:296-303."""
  sizes, size = [], max_size
  while True:
    sizes.append(size)
    if size <= state.min_res:
      break
    size = size // state.renderer_stride
  return sizes


def forward(state, inputs, training=True, build_loss=True):
  """IMMModel.build (imm_model.py:413-490).  Returns dict with loss, tensors, new buffer values."""
  im, future_im = inputs['image'], inputs['future_image']
  new_buffers = OrderedDict()
  R = future_im.shape[1]
  # image_encoder (imm_model.py:220-230): [im] + block features
  embeddings = [im] + encoder(state, new_buffers, 'model/image_encoder/encoder', im, training)
  # pose_encoder (imm_model.py:233-276)
  pf = encoder(state, new_buffers, 'model/pose_encoder/encoder', future_im, training)[-1]
  heatmaps = conv_block(state, new_buffers, 'model/pose_encoder', 'conv_1', pf, 1, False, False, training)
  gauss_y, gauss_y_prob = get_coord(heatmaps, 2, heatmaps.shape[1])
  gauss_x, gauss_x_prob = get_coord(heatmaps, 1, heatmaps.shape[2])
  gauss_mu = torch.stack([gauss_y, gauss_x], dim=2)
  sizes = render_sizes(state, R)
  pose_embeddings = [get_gaussian_maps(gauss_mu, [s, s], 1.0 / state.gauss_std, mode=state.gauss_mode) for s in sizes]
  # group by size; resize when missing (imm_model.py:311-335)
  grouped = {}
  for e in embeddings:
    grouped.setdefault(e.shape[1], []).append(e)
  for rs in sizes:
    if rs not in grouped:
      src = [s for s in sorted(grouped.keys()) if s >= rs][0]
      grouped[rs] = [resize_bilinear(e, [rs, rs], align_corners=True) for e in grouped[src]]
  gp = {}
  for e in pose_embeddings:
    gp.setdefault(e.shape[1], []).append(e)
  joint16 = torch.cat(grouped[16] + gp[16], dim=-1)          # only size 16 is consumed (imm_model.py:161)
  # simple_renderer (imm_model.py:154-179)
  n_out = 3 + (len(state.perceptual_comp) if state.channels_bug_fix else 0)
  x = joint16
  for name, _, bn, relu, up in renderer_spec(state.n_filters_render, R, n_out):
    x = conv_block(state, new_buffers, 'model/renderer', name, x, 1, bn, relu, training)
    if up:
      x = resize_bilinear(x, [2 * x.shape[1], 2 * x.shape[2]])
  future_im_pred = x[..., :3]                                  # imm_model.py:348-355
  out = {'future_im_pred': future_im_pred, 'gauss_yx': gauss_mu, 'heatmaps': heatmaps,
         'gauss_y_prob': gauss_y_prob, 'gauss_x_prob': gauss_x_prob,
         'pose_embedding_maps': pose_embeddings[0], 'joint16': joint16,
         'new_buffers': new_buffers, 'loss': None}
  if build_loss:
    mask = inputs.get('mask') if state.loss_mask else None
    rec, agg, levels = perceptual_loss(state, future_im, future_im_pred, mask, training)
    new_buffers.update(agg)
    wl = weight_decay_loss(state)
    out.update({'reconstruction_loss': rec, 'weights_loss': wl, 'loss': rec + wl, 'level_losses': levels})
  return out


def vgg_features(state, x_rgb, upto='conv5_2'):
  """build_vgg16 (build_vgg16.py:14-35) + vgg16.build_network (vgg16.py:289-375), batch_norm=False,
  activations post-ReLU (vgg16.py:229-236); net['input'] = raw RGB input (build_vgg16.py:34)."""
  net = OrderedDict()
  net['input'] = x_rgb
  x = x_rgb.mean(dim=3, keepdim=True)
  x = x / 255.0
  x = x - VGG_MEAN / 255.0
  for item in VGG_LAYERS:
    if isinstance(item, str):
      x = max_pool_2x2(x)
      net[item] = x
    else:
      name = item[0]
      w = state.vgg['SelfSupReconstructionLoss/vgg16/%s/weights' % name]
      b = state.vgg['SelfSupReconstructionLoss/vgg16/%s/biases' % name]
      x = torch.relu(conv2d_same(x, w, b, 1))
      net[name] = x
      if name == upto:
        break
  return net


def perceptual_loss(state, gt_image, pred_image, mask, training):
  """IMMModel._colorization_reconstruction_loss (imm_model.py:111-151), _loss_mask (:408-410),
  BaseModel._exp_running_avg (base_model.py:39-50; rho 0.99; NO stop-gradient on wl)."""
  ims = torch.cat([gt_image, pred_image], dim=0)
  net = vgg_features(state, ims)
  losses, new_agg, levels = [], OrderedDict(), []
  for k, nm in enumerate(state.perceptual_comp):
    f = net[nm]
    half = f.shape[0] // 2
    f_gt, f_pred = f[:half], f[half:]
    l = (f_gt - f_pred) ** 2
    if mask is not None:
      m = resize_bilinear(mask, [l.shape[1], l.shape[2]])
      masked = lambda t: t * m
    else:
      masked = lambda t: t
    s = masked(l).mean()
    a = state.buffers['SelfSupReconstructionLoss/%s_agg' % nm]
    wl = a + (1.0 - 0.99) * (s - a)
    if training:
      new_agg['SelfSupReconstructionLoss/%s_agg' % nm] = wl.detach()
    lk = masked(l / wl).mean()
    losses.append(lk)
    levels.append(lk.detach())
  return 1000.0 * sum(losses), new_agg, levels


def weight_decay_loss(state):
  """BaseModel._decay (base_model.py:33-37) over l2_regularizer(1e-5) terms (nn_utils.py:44-46):
  sum_w 1e-5 * 0.5 * ||w||^2, conv `w` tensors only."""
  tot = 0.0
  for k, v in state.params.items():
    if k.endswith('/w'):
      tot = tot + WD * 0.5 * (v * v).sum()
  return tot


def train_step(state, inputs, clip_value=1.0, lr_multiple=1.0, n_towers=1):
  """One optimisation step: train_single (cnn_train_multi.py:195-250) for n_towers == 1, train_multi
  (:109-192) otherwise: per-tower loss on an even batch split, tower-gradient MEAN, then per-tensor
  clip_by_norm, TF-Adam; BN moving stats / *_agg from the LAST tower (:155,166).
  Mutates `state`; returns dict(loss, grads (pre-clip, averaged), outputs of the last tower)."""
  for p in state.params.values():
    p.requires_grad_(True)
  B = inputs['image'].shape[0]
  assert B % n_towers == 0
  per = B // n_towers
  grads_sum, losses, out = None, [], None
  for t in range(n_towers):
    sub = {k: v[t * per:(t + 1) * per] for k, v in inputs.items()}
    out = forward(state, sub, training=True, build_loss=True)
    g = torch.autograd.grad(out['loss'], list(state.params.values()), allow_unused=True)
    g = [torch.zeros_like(p) if gi is None else gi for gi, p in zip(g, state.params.values())]
    grads_sum = g if grads_sum is None else [a + b for a, b in zip(grads_sum, g)]
    losses.append(out['loss'].detach())
  grads = OrderedDict((k, gs / n_towers) for k, gs in zip(state.params.keys(), grads_sum))
  lr = learning_rate(state.global_step, lr_multiple=lr_multiple)
  state.adam_t += 1
  with torch.no_grad():
    for k in list(state.params.keys()):
      g = clip_by_norm(grads[k], clip_value) if clip_value is not None else grads[k]
      var, m, v = adam_step(state.params[k].detach(), g, state.adam_m[k], state.adam_v[k], lr, state.adam_t)
      state.params[k] = var
      state.adam_m[k] = m
      state.adam_v[k] = v
    for k, v in out['new_buffers'].items():
      state.buffers[k] = v.detach()
  state.global_step += 1.0
  return {'loss': torch.stack(losses).mean(), 'grads': grads, 'out': out, 'lr': lr}


# ----------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------
def smooth_mask(h, w, margin=10, step=20, b=0.4, dtype=torch.float32):
  """TPSDataset._get_smooth_mask / _get_smooth_step (tps_dataset.py:47-67; margin 10, step 20:
  celeba_dataset.py:165), unwarped."""
  def smooth_step(n, bb):
    x = torch.linspace(-1.0, 1.0, n, dtype=torch.float32)
    return 0.5 + 0.5 * torch.tanh(x / bb)
  def strip(size):
    return torch.cat([torch.zeros(margin), smooth_step(step, b), torch.ones(size - 2 * margin - 2 * step),
                      smooth_step(step, -b), torch.zeros(margin)])
  return (strip(h)[:, None] * strip(w)[None]).to(dtype)


def synthetic_inputs(batch, image_size=128, seed=0, dtype=torch.float32):
  """Seeded smooth random image pairs in [0,255] + the reference's smooth border mask."""
  gen = torch.Generator().manual_seed(1000 + seed)
  def smooth_image():
    lo = torch.rand((batch, 3, image_size // 8, image_size // 8), generator=gen) * 255.0
    hi = F.interpolate(lo, size=(image_size, image_size), mode='bilinear', align_corners=False)
    return hi.permute(0, 2, 3, 1).contiguous()
  image = smooth_image()
  future = smooth_image()
  mask = smooth_mask(image_size, image_size).view(1, image_size, image_size, 1).repeat(batch, 1, 1, 1)
  return {'image': image.to(dtype), 'future_image': future.to(dtype), 'mask': mask.to(dtype)}
