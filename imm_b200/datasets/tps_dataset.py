"""TPSDataset -- GPU mirror of imm/datasets/tps_dataset.py: turns a batch of single images into
(image, future_image, mask) training pairs with two random thin-plate-spline warps, entirely on the device.

File listing / JPEG decoding of the real datasets stays out of scope (celeba/aflw are not reachable offline);
`sample_base_images` is the hook a real reader plugs into: it must return [B,R,R,3] fp32 images in [0,255]."""
import numpy as np
import torch

from ..utils.synthetic import smooth_mask, synthetic_inputs
from ..utils.tps_sampler import TPSRandomSampler, apply_tps


class TPSDataset(object):
  """Same constructor defaults as the reference (tps_dataset.py:17-45)."""

  def __init__(self, data_dir=None, subset='train', max_samples=None, image_size=[128, 128], order_stream=False,
               landmarks=False, tps=True, vertical_points=10, horizontal_points=10, rotsd=[0.0, 5.0],
               scalesd=[0.0, 0.1], transsd=[0.1, 0.1], warpsd=[0.001, 0.005, 0.001, 0.01], name='TPSDataset',
               device='cuda:0', seed=0):
    if landmarks and tps:
      raise ValueError('Outputing landmarks is not supported with TPS transform.')      # tps_dataset.py:27-28
    self._image_size, self._tps, self._device, self._seed = image_size, tps, device, seed
    rng = np.random.RandomState(seed)
    if tps:
      self._target_sampler = TPSRandomSampler(image_size[1], image_size[0], vertical_points, horizontal_points,
                                              rotsd=rotsd[0], scalesd=scalesd[0], transsd=transsd[0],
                                              warpsd=warpsd[:2], pad=False, device=device, rng=rng)
      self._source_sampler = TPSRandomSampler(image_size[1], image_size[0], vertical_points, horizontal_points,
                                              rotsd=rotsd[1], scalesd=scalesd[1], transsd=transsd[1],
                                              warpsd=warpsd[2:], pad=False, device=device, rng=rng)

  def _get_smooth_mask(self, h, w, margin, step):
    """tps_dataset.py:53-67."""
    return smooth_mask(h, w, margin, step)

  def sample_base_images(self, batch_size, i):
    """Stand-in for the JPEG reader: seeded smooth random images."""
    return synthetic_inputs(batch_size, self._image_size[0], seed=self._seed + i)['image']

  def get_dataset(self, batch_size, repeat=True, shuffle=False, num_preprocess_threads=12, rank=0):
    R = self._image_size[0]
    mask = self._get_smooth_mask(R, R, 10, 20).view(1, R, R, 1).repeat(batch_size, 1, 1, 1).to(self._device)   # celeba_dataset.py:165
    state = {'i': 0}

    def next_batch():
      img = self.sample_base_images(batch_size, 1000 * rank + state['i']).to(self._device, non_blocking=True)
      state['i'] += 1
      if not self._tps:
        return {'image': img, 'future_image': img, 'mask': mask}
      return apply_tps(img, mask, self._target_sampler, self._source_sampler)      # tps_dataset.py:70-96
    return next_batch
