"""Training summaries -- TensorBoard event files with the reference's tags:
image summaries future_im / im / pose_embedding / future_im_pred / mask (imm_model.py:21-24,456-468), raw +
moving-average cost scalars (base_model.py:52-60), learning rate (scripts/train.py:111).
Host-side only (runs every `n_summary` steps, off the hot path); device tensors are read back here."""
import random

import numpy as np
import torch


def get_random_color(pastel_factor=0.5, rnd=random):
  """imm/utils/utils.py:236-237."""
  return [(x + pastel_factor) / (1.0 + pastel_factor) for x in [rnd.uniform(0, 1.0) for _ in [1, 2, 3]]]


def color_distance(c1, c2):
  return sum([abs(x[0] - x[1]) for x in zip(c1, c2)])


def generate_new_color(existing_colors, pastel_factor=0.5, rnd=random):
  """imm/utils/utils.py:244-256: best of 100 random candidates by minimum L1 distance to the existing colours."""
  max_distance, best_color = None, None
  for _ in range(0, 100):
    color = get_random_color(pastel_factor=pastel_factor, rnd=rnd)
    if not existing_colors:
      return color
    best_distance = min([color_distance(color, c) for c in existing_colors])
    if not max_distance or best_distance > max_distance:
      max_distance, best_color = best_distance, color
  return best_color


def get_n_colors(n, pastel_factor=0.9, rnd=random):
  """imm/utils/utils.py:259-263 (NB the reference ignores its argument and always uses 0.9)."""
  colors = []
  for _ in range(n):
    colors.append(generate_new_color(colors, pastel_factor=0.9, rnd=rnd))
  return colors


def colorize_landmark_maps(maps, colors=None):
  """imm_model.py:81-91: [B,H,W,N] landmark maps -> [B,H,W,3], each landmark in its own colour, max over landmarks."""
  n_maps = maps.shape[-1]
  if colors is None:
    colors = get_n_colors(n_maps, pastel_factor=0.0)
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

  def write(self, model, step, lr=None):
    eng = model.engine
    self._img('future_im', eng.future_image, step)                        # imm_model.py:456-457
    self._img('im', eng.image, step)
    if self._colors is None:
      self._colors = get_n_colors(eng.K, pastel_factor=0.0)
    maps = eng.gaussian_maps(eng.mu[:self.max_outputs], eng.R)           # pose_embeddings[0]: full-resolution maps
    self._img('pose_embedding', colorize_landmark_maps(maps, self._colors) * 255.0, step)
    self._img('future_im_pred', eng.pred[..., :3], step)                  # clipped to [0,255] (:467)
    if eng.mask is not None:
      self._img('mask', eng.mask * 255.0, step)
    for op in model._avg_ops:                                             # cost EMAs (base_model.py:52-60)
      op()
    for name, v in model._cost_raw.items():
      self.writer.add_scalar('train/%s_raw' % name, v, step)
      self.writer.add_scalar('train/%s_avg' % name, model._cost_avgs[name], step)
    if lr is not None:
      self.writer.add_scalar('lr', lr, step)
    self.writer.flush()
