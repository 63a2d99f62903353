"""Class for IMM models -- host-side mirror of imm/models/imm_model.py (the drop-in boundary).

Same class / method / argument names as the reference; graph-construction arguments (`costs_collection`,
`scope`, `var_device`) are accepted and ignored because the work is executed eagerly by hand-written sm_100a
kernels (imm_b200/engine.py -> libimm_b200.so).  file:line citations are under /root/reference."""
import torch

from .. import _lib
from ..engine import IMMEngine
from ..models.base_model import BaseModel


def get_gaussian_maps(mu, shape_hw, inv_std, mode='ankush'):
  """imm_model.py:34-78.  mu [B,NMAPS,2] (y,x) -> [B,SHAPE_H,SHAPE_W,NMAPS] on mu's device.
  Only mode 'rot' (used by every shipped config) has a CUDA kernel."""
  if mode != 'rot':
    raise ValueError("mode %r is not built on the CUDA path (configs use 'rot'): " % mode + str(mode))
  assert shape_hw[0] == shape_hw[1]
  if not mu.is_cuda:
    raise _lib.ImmbError('get_gaussian_maps: mu must be a CUDA tensor (no CPU fallback)')
  B, K = mu.shape[0], mu.shape[1]
  out = torch.empty((B, shape_hw[0], shape_hw[1], K), dtype=torch.float32, device=mu.device)
  _lib.call('immb_gaussian_maps', mu.contiguous().float(), B, K, int(shape_hw[0]), float(inv_std), out,
            _lib.stream_ptr())
  return out


class IMMModel(BaseModel):
  """IMMModel(config, global_step=None, dtype=float32, name='IMMModel')   (imm_model.py:95-101)

  config: the `model` section of the reference's experiment YAMLs (attribute access + hasattr probing).
  Extra keyword-only knobs select the device-side execution: `device`, `precision`, `engine`, `world_size`."""

  def __init__(self, config, global_step=None, dtype=torch.float32, name='IMMModel', device='cuda:0',
               precision=None, engine=_lib.ENGINE_AUTO, world_size=1, vgg_data=None, seed=0):
    super(IMMModel, self).__init__(dtype, name)
    self._config = config
    self._global_step = global_step
    self._device, self._precision, self._engine_sel, self._world = device, precision, engine, world_size
    self._vgg_data, self._seed = vgg_data, seed
    self.engine = None
    self._colors = None
    self._tensors = {}            # the reference's 'tensors' collection (imm_model.py:250,266-267)

  # -- construction -----------------------------------------------------------------------------------
  def _ensure_engine(self, batch, image_size):
    if self.engine is not None:
      if (self.engine.B, self.engine.R) != (batch, image_size):
        raise _lib.ImmbError('IMMModel was built for batch %d / size %d; got %d / %d (static shapes, like the '
                             'reference graph)' % (self.engine.B, self.engine.R, batch, image_size))
      return self.engine
    eng = IMMEngine(self._config, batch, image_size, self._device, self._precision, self._engine_sel, self._world)
    eng.init_parameters(self._seed)
    if self._global_step is not None:
      eng.global_step = float(self._global_step)
    self.engine = eng
    if self._vgg_data is not None:
      eng.load_vgg_caffe_dict(self._vgg_data)
    return eng

  def load_vgg(self, data=None):
    """build_vgg16 loads `config.perceptual.net_file` with deepdish (build_vgg16.py:16); offline that file does
    not exist, so the caller passes the Caffe-layout dict (real or imm_b200.utils.synthetic)."""
    if data is None:
      from ..utils.vgg_io import load_caffe_h5
      data = load_caffe_h5(self._config.perceptual.net_file)
    self._vgg_data = data
    if self.engine is not None:
      self.engine.load_vgg_caffe_dict(data)

  @staticmethod
  def _to_device(t, dev):
    if t.is_cuda:
      return t.contiguous()
    return t.to(dev, non_blocking=True).contiguous()

  # -- sub-builders: each runs the engine section of the same name on its own ---------------------------------
  # They exist for callers that assemble the model piecewise (imm_model.py:279-357 `model` takes them as callables).
  # The engine has static shapes and one set of buffers: outputs are views / fresh fp32 copies of engine buffers and
  # stay valid until the next call.  On an fp16-plane engine the activation exponents must have been calibrated by one
  # full build() in the same training_pl mode before a section is run on its own.
  def _section_engine(self, x, training_pl):
    B, R = int(x.shape[0]), int(x.shape[1])
    eng = self._ensure_engine(B, R) if x.shape[-1] == 3 else self.engine
    if eng is None:
      raise _lib.ImmbError('run build() once before calling a sub-builder on intermediate features')
    if eng.h16 and eng._scale_mode is None:
      raise _lib.ImmbError('fp16-plane engine: run one full build() (same training_pl) before a stand-alone sub-builder, '
                           'so that the activation exponents are calibrated')
    return eng

  @staticmethod
  def _block_outputs(layers):
    """The four block outputs of an encoder (after conv_2 / conv_4 / conv_6 / conv_8, imm_model.py:195-216), fp32."""
    return [L.out.value()[..., :L.cout] for L in layers[1::2]]

  def encoder(self, x, training_pl, var_device='/cpu:0', _branch='image_encoder'):
    """imm_model.py:182-217 -> [f(R), f(R/2), f(R/4), f(R/8)] block features of x [B,R,R,3]."""
    eng = self._section_engine(x, training_pl)
    eng.run_encoder(_branch, self._to_device(x, eng.dev), bool(training_pl))
    return self._block_outputs(eng.enc_layers[_branch])

  def image_encoder(self, x, training_pl, filters=64, var_device='/cpu:0'):
    """imm_model.py:220-230: [x] + encoder(x)."""
    eng = self._section_engine(x, training_pl)
    eng.run_image_encoder(self._to_device(x, eng.dev), bool(training_pl))
    return [x] + self._block_outputs(eng.enc_layers['image_encoder'])

  def pose_encoder(self, x, training_pl, n_maps=1, filters=32, gauss_mode='ankush', map_sizes=None, reuse=False,
                   var_device='/cpu:0'):
    """imm_model.py:233-276 -> (gauss_mu [B,K,2] (y,x) in [-1,1], [Gaussian maps per size in map_sizes])."""
    eng = self._section_engine(x, training_pl)
    if n_maps not in (1, eng.K) or gauss_mode not in ('ankush', 'rot'):
      raise ValueError('the engine was built for n_maps=%d, gauss_mode=rot' % eng.K)
    eng.run_pose_branch(self._to_device(x, eng.dev), bool(training_pl))
    self._tensors = {'heatmaps': eng.pose_conv.y[..., :eng.K], 'gauss_y_prob': eng.py, 'gauss_x_prob': eng.px}
    maps = [get_gaussian_maps(eng.mu, [sz, sz], eng.inv_std, 'rot') for sz in (map_sizes or [])]
    return eng.mu, maps

  def simple_renderer(self, feat_heirarchy, training_pl, n_final_out=3, final_res=128, var_device='/cpu:0'):
    """imm_model.py:154-179: renders feat_heirarchy[16] = [B,16,16,enc_feat+K] -> [B,final_res,final_res,n_final_out]."""
    eng = self._section_engine(feat_heirarchy[16], training_pl)
    if final_res != eng.R or n_final_out > eng.n_out:
      raise ValueError('the engine renders %dx%d images with %d channels' % (eng.R, eng.R, eng.n_out))
    eng.load_joint(feat_heirarchy[16])
    return eng.run_renderer(bool(training_pl))[..., :n_final_out]

  def model(self, im, future_im, image_encoder=None, pose_encoder=None, renderer=None):
    """imm_model.py:279-357 -> (future_im_pred [B,R,R,3], gauss_yx [B,K,2], [pose embedding maps per render size]).
    The three callables of the reference signature are accepted; the wiring (concat, optional align_corners resize,
    first three renderer channels) is the engine's forward pass."""
    B, R = int(future_im.shape[0]), int(future_im.shape[1])
    eng = self._ensure_engine(B, R)
    training = bool(self._opts['training_pl']) if self._opts else False
    eng.forward(self._to_device(im, eng.dev), self._to_device(future_im, eng.dev), None, training=training,
                build_loss=False)
    sizes, size = [], R
    while size >= int(self._config.min_res):            # imm_model.py:296-303
      sizes.append(size)
      size //= int(self._config.renderer_stride)
    return eng.pred[..., :3], eng.mu, [get_gaussian_maps(eng.mu, [sz, sz], eng.inv_std, 'rot') for sz in sizes]

  def loss(self, future_im_pred, future_im, future_yx, future_yx_gmaps, costs_collection, training_pl,
           loss_mask=None):
    """imm_model.py:360-405 on the engine's current prediction (future_im_pred must be the tensor model() / build()
    returned): reconstruction + weight decay, cost summaries registered once."""
    eng = self.engine
    if eng is None or future_im_pred.data_ptr() != eng.pred.data_ptr():
      raise _lib.ImmbError('loss(): future_im_pred must be the prediction the engine just produced')
    if bool(self._config.loss_mask) and loss_mask is None:
      raise RuntimeError('No loss mask recieved but is required.')
    eng.future_image = self._to_device(future_im, eng.dev)
    eng.mask = self._to_device(loss_mask, eng.dev) if (eng.use_mask and loss_mask is not None) else None
    eng._gt_tower_forked = False
    eng.fwd_pool[-len(eng.comp):].zero_()
    eng._loss_fwd(bool(training_pl))
    return eng.loss_value().view(())

  def _decay(self, scope=None):
    """base_model.py:33-37: the sum of the L2 weight-decay terms (device scalar; refreshed by loss_value())."""
    if self.engine is None:
      raise _lib.ImmbError('_decay(): the model has not been built')
    self.engine.loss_value()
    return self.engine.weights_loss.view(())

  def _loss_mask(self, map, mask):
    """imm_model.py:408-410: map * resize_images(mask, map.shape) -- at the integer scales used by the loss the TF1
    legacy bilinear resize is a pure subsample, which the loss kernels index directly."""
    s = mask.shape[1] // map.shape[1]
    return map * mask[:, ::s, ::s]

  # -- the public entry point ----------------------------------------------------------------------------
  def build(self, inputs, training_pl, costs_collection='costs', scope=None, var_device='/cpu:0',
            output_tensors=False, build_loss=True):
    """imm_model.py:413-490.  inputs: dict with 'image', 'future_image' [B,R,R,3] fp32 in [0,255] and (when
    config.loss_mask) 'mask' [B,R,R,1]; host (ideally pinned) or CUDA tensors.  training_pl: bool.
    Returns (None, loss, avg_ops[, tensors]) like the reference; `loss` is a CUDA scalar tensor."""
    im, future_im = inputs['image'], inputs['future_image']
    B, R = int(future_im.shape[0]), int(future_im.shape[1])
    assert future_im.shape[1] == future_im.shape[2]
    eng = self._ensure_engine(B, R)
    dev = eng.dev
    im_d, fut_d = self._to_device(im, dev), self._to_device(future_im, dev)
    mask_d = self._to_device(inputs['mask'], dev) if 'mask' in inputs else None
    if build_loss and bool(self._config.loss_mask) and mask_d is None:
      raise RuntimeError('No loss mask recieved but is required.')     # imm_model.py:363-367
    training = bool(training_pl)
    eng.forward(im_d, fut_d, mask_d, training=training, build_loss=build_loss)
    loss = None
    if build_loss:
      loss = eng.loss_value().view(())
      if not self._avg_ops:
        self._add_cost_summary(lambda: eng.rec_loss.item(), 'reconstruction_loss')     # imm_model.py:390
        self._add_cost_summary(lambda: eng.weights_loss.item(), 'weights_loss')        # :396
        self._add_cost_summary(lambda: eng.total_loss.item(), 'loss_total')            # :402
    self._tensors = {'heatmaps': eng.pose_conv.y[..., :eng.K], 'gauss_y_prob': eng.py, 'gauss_x_prob': eng.px}
    if output_tensors:
      tensors = {}
      tensors.update(inputs)
      from ..utils.summaries import colorize_landmark_maps, get_n_colors
      if self._colors is None:
        self._colors = get_n_colors(eng.K)
      pose_embedding = colorize_landmark_maps(get_gaussian_maps(eng.mu, [R, R], eng.inv_std, 'rot'), self._colors)
      tensors.update({'future_im': fut_d, 'im': im_d,
                      'pose_embedding': pose_embedding,                     # imm_model.py:463-465,480-485
                      'future_im_pred': eng.pred[..., :3],
                      'gauss_yx': eng.mu})
      return None, loss, self._avg_ops, tensors
    return None, loss, self._avg_ops

  def train_step(self, inputs, clip_value=None, lr=None, beta1=0.9, beta2=0.999, eps=1e-8, allreduce=None):
    """build(inputs, training_pl=True) + backward + clip/Adam as ONE engine call (what session.run(train_op) executes,
    cnn_train_multi.py:445-460), so that the engine can replay it as a CUDA graph.  Returns the loss (CUDA scalar)."""
    im, future_im = inputs['image'], inputs['future_image']
    B, R = int(future_im.shape[0]), int(future_im.shape[1])
    eng = self._ensure_engine(B, R)
    dev = eng.dev
    im_d, fut_d = self._to_device(im, dev), self._to_device(future_im, dev)
    mask_d = self._to_device(inputs['mask'], dev) if 'mask' in inputs else None
    if bool(self._config.loss_mask) and mask_d is None:
      raise RuntimeError('No loss mask recieved but is required.')     # imm_model.py:363-367
    loss = eng.train_step(im_d, fut_d, mask_d, clip_value=clip_value, lr=lr, beta1=beta1, beta2=beta2, eps=eps,
                          allreduce=allreduce)
    if not self._avg_ops:
      self._add_cost_summary(lambda: eng.rec_loss.item(), 'reconstruction_loss')     # imm_model.py:390
      self._add_cost_summary(lambda: eng.weights_loss.item(), 'weights_loss')        # :396
      self._add_cost_summary(lambda: eng.total_loss.item(), 'loss_total')            # :402
    self._tensors = {'heatmaps': eng.pose_conv.y[..., :eng.K], 'gauss_y_prob': eng.py, 'gauss_x_prob': eng.px}
    return loss.view(())

  def get_collection(self, name='tensors'):
    return dict(self._tensors)

  # -- checkpoint layout: TF variable names -> tensors -------------------------------------------
  def state_dict(self, include_optimizer=True):
    eng = self.engine
    sd = {}
    for d in (eng.params, eng.buffers, eng.vgg_params if eng.vgg_loaded else {}):
      for k, v in d.items():
        sd[k] = v.detach().cpu().clone()
    sd['global_step'] = torch.tensor(eng.global_step)
    if include_optimizer:
      b1, b2 = eng.adam_betas                                          # the betas of the optimiser that ran the steps
      sd['beta1_power'] = torch.tensor(b1 ** (eng.adam_t + 1))        # TF keeps beta^t for the NEXT step
      sd['beta2_power'] = torch.tensor(b2 ** (eng.adam_t + 1))
      sd['__adam_t'] = torch.tensor(eng.adam_t)
      for k in eng.params:
        sd[k + '/Adam'] = eng.adam_m[k].detach().cpu().clone()
        sd[k + '/Adam_1'] = eng.adam_v[k].detach().cpu().clone()
    return sd

  def save_checkpoint(self, prefix, include_optimizer=True):
    """tf.train.Saver(tf.global_variables()).save (cnn_train_multi.py:439,511-513): writes `<prefix>.index` and
    `<prefix>.data-00000-of-00001` in TensorFlow's TensorBundle format (imm_b200/utils/tf_checkpoint.py), variables
    under the reference's names, float32, HWIO weights.  `__adam_t` is not a TF variable and is not written: the Adam
    step count is recovered from `beta1_power` on restore."""
    from ..utils import tf_checkpoint
    sd = self.state_dict(include_optimizer)
    sd.pop('__adam_t', None)
    return tf_checkpoint.write_checkpoint(prefix, {k: v.numpy().astype('float32') for k, v in sd.items()})

  def restore_checkpoint(self, fname, **kwargs):
    """tf.train.Saver(var_list).restore(session, fname) (cnn_train_multi.py:404-433): fname is a TensorBundle prefix
    (`model.ckpt-N`, with `.index` next to it) or a legacy torch.save file of state_dict()."""
    import os
    from ..utils import tf_checkpoint
    if os.path.exists(fname + '.index'):
      # only what this model can use is decoded: a checkpoint may carry entries without a numeric encoding here
      # (DT_STRING object graphs of later TF1 savers, partitioned variables) that are none of this model's business
      eng = self.engine
      wanted = set(eng.params) | set(eng.buffers) | {'global_step', 'beta1_power', 'beta2_power'}
      wanted |= {k + s_ for k in eng.params for s_ in ('/Adam', '/Adam_1')}
      reader = tf_checkpoint.CheckpointReader(fname)
      sd = {k: torch.from_numpy(reader.get_tensor(k)) for k in reader.entries
            if k in wanted or k.startswith('SelfSupReconstructionLoss/vgg16/')}
    elif os.path.exists(fname):
      sd = torch.load(fname, map_location='cpu')
    else:
      raise IOError('model file does not exist at: ' + fname)
    self.load_state_dict(sd, **kwargs)
    return sd

  def load_state_dict(self, sd, vars_to_restore='model', ignore_missing_vars=False, reset_global_step=-1,
                      exclude_vars=None, adam_betas=None):
    """cnn_train_multi.py:404-433 semantics: 'model' = MODEL_VARIABLES (w, b, *_agg, global_step -- NOT the
    tf.layers BN variables), 'all' = every global variable incl. BN and Adam slots (--restore-optim)."""
    eng = self.engine
    model_vars = [k for k in eng.params if k.endswith('/w') or k.endswith('/b')]
    model_vars += [k for k in eng.buffers if k.endswith('_agg')]
    all_vars = list(eng.params.keys()) + list(eng.buffers.keys())
    names = list(all_vars if vars_to_restore == 'all' else model_vars)
    for ex in (exclude_vars or []):        # cnn_train_multi.py:425-430: drops the FIRST variable whose name contains ex
      hit = [i for i, n in enumerate(names) if ex in n]
      if hit:
        names.pop(hit[0])
    missing = [k for k in names if k not in sd]
    if vars_to_restore == 'all':     # tf.global_variables() includes the optimiser slots: the reference's Saver fails without them
      missing += [k + s_ for k in eng.params for s_ in ('/Adam', '/Adam_1') if k + s_ not in sd]
      missing += [k for k in ('beta1_power', 'beta2_power') if k not in sd and '__adam_t' not in sd]
    if missing and not ignore_missing_vars:
      raise KeyError('variables missing from the checkpoint: %s' % missing[:5])
    params = {k: sd[k] for k in names if k in sd and k in eng.params}
    buffers = {k: sd[k] for k in names if k in sd and k in eng.buffers}
    adam_m = adam_v = None
    if vars_to_restore == 'all':
      adam_m = {k: sd[k + '/Adam'] for k in eng.params if k + '/Adam' in sd}
      adam_v = {k: sd[k + '/Adam_1'] for k in eng.params if k + '/Adam_1' in sd}
      if '__adam_t' in sd:
        eng.adam_t = int(sd['__adam_t'])
      elif 'beta1_power' in sd:       # TF keeps beta1^(t+1) (AdamOptimizer._finish); t = steps applied so far
        import math
        b1, b2 = adam_betas if adam_betas is not None else eng.adam_betas      # betas of the run that wrote the file
        b1p, b2p = float(sd['beta1_power']), float(sd.get('beta2_power', 0.0))
        if b1p > 1e-30 and 0.0 < b1 < 1.0:
          eng.adam_t = max(int(round(math.log(b1p) / math.log(b1))) - 1, 0)
        elif b2p > 1e-30 and 0.0 < b2 < 1.0:
          eng.adam_t = max(int(round(math.log(b2p) / math.log(b2))) - 1, 0)
        else:                         # both powers underflowed in fp32 (> ~87k steps): the bias correction is 1 anyway
          eng.adam_t = max(int(float(sd.get('global_step', 0))) + 1, 100000)
    eng.load_state(params, buffers, adam_m, adam_v)
    if vars_to_restore == 'all':
      # the frozen tower's weights/biases are global variables too: a full checkpoint (the reference's own included)
      # carries them, so a run can start from a checkpoint alone when the Caffe HDF5 file is not at hand
      np_vars = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in sd.items()
                 if k.startswith('SelfSupReconstructionLoss/vgg16/')}
      if np_vars:
        eng.load_vgg_hwio(np_vars)
    if reset_global_step >= 0:
      eng.global_step = float(reset_global_step)
    elif 'global_step' in sd:
      eng.global_step = float(sd['global_step'])
