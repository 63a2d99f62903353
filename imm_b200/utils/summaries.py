"""Training summaries -- TensorBoard event files with the reference's tags:
image summaries future_im / im / pose_embedding / future_im_pred / mask (imm_model.py:21-24,456-468), raw +
moving-average cost scalars (base_model.py:52-60), learning rate (scripts/train.py:111).
Host-side only (runs every `n_summary` steps, off the hot path); device tensors are read back here."""
import colorsys

import numpy as np
import torch


def get_n_colors(n, pastel_factor=0.9, rnd=None):
  """N visually distinct RGB colours in [pastel/(1+pastel), 1]^3, the range of the reference's palette
  (imm/utils/utils.py:259-263 draws "maximally different" random pastel colours; which colours come out is not part
  of any contract -- they only tint the `pose_embedding` image summary, imm_model.py:81-91,463-465).
  Here: golden-ratio hue stepping at full saturation / value (consecutive hues are >= 0.38 of the circle apart),
  then the same pastel blend (c + p) / (1 + p).  `rnd` (a random.Random) only picks the starting hue."""
  h0 = rnd.random() if rnd is not None else 0.0
  out = []
  for i in range(n):
    r, g, b = colorsys.hsv_to_rgb((h0 + i * 0.6180339887498949) % 1.0, 1.0, 1.0)
    out.append([(c + pastel_factor) / (1.0 + pastel_factor) for c in (r, g, b)])
  return out


def colorize_landmark_maps(maps, colors=None):
  """imm_model.py:81-91: [B,H,W,N] landmark maps -> [B,H,W,3], each landmark in its own colour, max over landmarks."""
  n_maps = maps.shape[-1]
  if colors is None:
    colors = get_n_colors(n_maps)
  col = torch.as_tensor(np.asarray(colors, dtype=np.float32), device=maps.device)      # [N,3]
  return (maps.unsqueeze(-1) * col.view(1, 1, 1, n_maps, 3)).amax(dim=3)


class SummaryLogger(object):
  """Writes the reference's train summaries for an IMMModel whose engine just ran a training forward pass."""

  def __init__(self, log_dir, max_outputs=1):
    from torch.utils.tensorboard import SummaryWriter
    self.writer = SummaryWriter(log_dir)
    self.max_outputs = max_outputs
    self._colors = None

  def _img(self, tag, t, step):
    x = t[:self.max_outputs].detach().float().clamp(0, 255).cpu() / 255.0       # NHWC
    if x.shape[-1] == 1:
      x = x.repeat(1, 1, 1, 3)
    self.writer.add_images('train/' + tag, x, step, dataformats='NHWC')

  def write(self, model, step, lr=None, advance_avgs=False):
    eng = model.engine
    self._img('future_im', eng.future_image, step)                        # imm_model.py:456-457
    self._img('im', eng.image, step)
    if self._colors is None:
      self._colors = get_n_colors(eng.K)
    maps = eng.gaussian_maps(eng.mu[:self.max_outputs], eng.R)           # pose_embeddings[0]: full-resolution maps
    self._img('pose_embedding', colorize_landmark_maps(maps, self._colors) * 255.0, step)
    self._img('future_im_pred', eng.pred[..., :3], step)                  # clipped to [0,255] (:467)
    if eng.mask is not None:
      self._img('mask', eng.mask * 255.0, step)
    if advance_avgs:                 # train_loop advances the cost EMAs every step (base_model.py:52-60); only a
      for op in model._avg_ops:      # caller that does not (fwd_only timing runs) asks for it here
        op()
    for name, v in model._cost_raw.items():
      self.writer.add_scalar('train/%s_raw' % name, v, step)
      self.writer.add_scalar('train/%s_avg' % name, model._cost_avgs[name], step)
    if lr is not None:
      self.writer.add_scalar('lr', lr, step)
    self.writer.flush()
