"""Random thin-plate-spline warps on the GPU -- host-side mirror of imm/utils/tps_sampler.py.

The reference generates the warp pairs with PyTorch-0.4 on the CPU through tf.py_func, serialised
(tps_dataset.py:82-83,155): its real-data input bottleneck.  Here the TPS grid evaluation and the bilinear
resampling are one CUDA kernel (immb_tps_warp); only the (Hc*Wc+3) x 2 random parameters are drawn on the host
(numpy, same draw order as sample_tps_w, tps_sampler.py:168-189)."""
import random

import numpy as np
import torch

from .. import _lib


def sample_tps_w(Hc, Wc, warpsd, rotsd, scalesd, transsd, rng=np.random):
  """Randomly sampled TPS params [(Hc*Wc+3), 2] (tps_sampler.py:168-189)."""
  Nc = Hc * Wc
  mask = (rng.rand(Nc, 2) > 0.5).astype(np.float32)
  W = warpsd[0] * rng.randn(Nc, 2) + warpsd[1] * (mask * rng.randn(Nc, 2))
  rnd = rng.randn
  rot = np.deg2rad(rnd() * rotsd)
  sc = 1.0 + rnd() * scalesd
  aff = [[transsd * rnd(), transsd * rnd()], [sc * np.cos(rot), sc * -np.sin(rot)], [sc * np.sin(rot), sc * np.cos(rot)]]
  return np.r_[W, aff]


class TPSRandomSampler(object):
  """TPSRandomSampler(height, width, vertical_points=10, horizontal_points=10, rotsd, scalesd, transsd, warpsd,
  cache_size=1000, cache_evict_prob=0.01, pad=True, device=None)  (tps_sampler.py:12-75).
  The cache holds parameter sets (equivalent to the reference's cache of evaluated grids)."""

  def __init__(self, height, width, vertical_points=10, horizontal_points=10, rotsd=0.0, scalesd=0.0, transsd=0.1,
               warpsd=(0.001, 0.005), cache_size=1000, cache_evict_prob=0.01, pad=True, device='cuda:0', rng=None):
    if pad:
      raise NotImplementedError('pad=True (replicate-pad by half the size) is not used by the datasets '
                                '(tps_dataset.py:37-45 pass pad=False) and is not built')
    self.input_height, self.input_width = height, width
    self.vertical_points, self.horizontal_points = vertical_points, horizontal_points
    self.rotsd, self.scalesd, self.transsd, self.warpsd = rotsd, scalesd, transsd, warpsd
    self.cache_size, self.cache_evict_prob = cache_size, cache_evict_prob
    self.cache = [None] * cache_size
    self.device = torch.device(device)
    self.rng = rng if rng is not None else np.random

  def _sample_w(self):
    return sample_tps_w(self.vertical_points, self.horizontal_points, self.warpsd, self.rotsd, self.scalesd,
                        self.transsd, self.rng).astype(np.float32)

  def _get_params(self, batch_size):
    ws = []
    for _ in range(batch_size):                      # tps_sampler.py:58-72
      entry = random.randint(0, self.cache_size - 1)
      if self.cache[entry] is None or random.random() < self.cache_evict_prob:
        self.cache[entry] = self._sample_w()
      ws.append(self.cache[entry])
    return torch.from_numpy(np.stack(ws)).to(self.device, non_blocking=True)

  def warp(self, x_nhwc, w_tps):
    """x [B,H,W,C] CUDA fp32, w_tps [B, M+3, 2] -> warped [B,H,W,C]."""
    B, H, W, C = x_nhwc.shape
    out = torch.empty_like(x_nhwc)
    _lib.call('immb_tps_warp', x_nhwc.contiguous(), B, H, W, C, w_tps.contiguous().float(), self.vertical_points,
              self.horizontal_points, out, _lib.stream_ptr())
    return out

  def forward(self, input_nhwc):
    return self.warp(input_nhwc, self._get_params(input_nhwc.shape[0]))

  def forward_py(self, input):
    """numpy NHWC in / out, like the reference's py_func entry (tps_sampler.py:101-108)."""
    x = torch.from_numpy(np.ascontiguousarray(input, dtype=np.float32)).to(self.device)
    return self.forward(x).cpu().numpy()


def apply_tps(image, mask, target_sampler, source_sampler):
  """TPSDataset._apply_tps (tps_dataset.py:70-96) on the GPU: returns the model's `inputs` dict."""
  x = torch.cat([mask, image], dim=3).contiguous()
  fut = target_sampler.forward(x)
  src = source_sampler.forward(fut)
  return {'image': src[..., 1:].contiguous(), 'future_image': fut[..., 1:].contiguous(),
          'mask': fut[..., 0:1].contiguous()}
