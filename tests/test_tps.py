"""GPU thin-plate-spline pair generator vs the restated reference (oracle/tps_oracle.py)."""
import numpy as np
import pytest
import torch

from oracle import tps_oracle as T
from oracle import imm_oracle as O


def test_tps_oracle_identity_and_translation_kat():
  """Hand-derived: zero non-linear part + identity affine rows = identity warp; affine offset (tx, 0) with
  corner-aligned coordinates shifts the sampling position by tx*(W-1)/2 pixels."""
  H = W = 9
  x = torch.arange(H * W, dtype=torch.float32).view(1, H, W, 1)
  w = torch.zeros(1, 103, 2)
  w[0, 101, 0] = 1.0          # x' = x
  w[0, 102, 1] = 1.0          # y' = y
  np.testing.assert_allclose(T.tps_warp_nhwc(x, w).numpy(), x.numpy(), atol=1e-4)
  w[0, 100, 0] = 2.0 / (W - 1)          # +1 pixel in x
  y = T.tps_warp_nhwc(x, w)[0, :, :, 0]
  np.testing.assert_allclose(y[:, :-1].numpy(), x[0, :, 1:, 0].numpy(), atol=1e-3)
  np.testing.assert_allclose(y[:, -1].numpy(), 0.0, atol=1e-3)       # zero padding outside the image


def test_sample_tps_w_matches_reference_draw_order():
  from imm_b200.utils.tps_sampler import sample_tps_w
  a = T.sample_tps_w(10, 10, (0.001, 0.005), 5.0, 0.05, 0.1, np.random.RandomState(3))
  b = sample_tps_w(10, 10, (0.001, 0.005), 5.0, 0.05, 0.1, np.random.RandomState(3))
  assert a.shape == (103, 2) and np.array_equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(2, 128, 128, 4), (3, 40, 56, 1)])
def test_tps_warp_kernel_matches_oracle(shape):
  from imm_b200.utils.tps_sampler import TPSRandomSampler
  B, H, W, C = shape
  rng = np.random.RandomState(5)
  x = torch.rand(B, H, W, C) * 255
  w = torch.from_numpy(np.stack([T.sample_tps_w(10, 10, (0.001, 0.005), 10.0, 0.05, 0.05, rng) for _ in range(B)]).astype(np.float32))
  ref = T.tps_warp_nhwc(x, w)
  s = TPSRandomSampler(H, W, rotsd=10.0, scalesd=0.05, transsd=0.05, pad=False)
  out = s.warp(x.cuda(), w.cuda()).cpu()
  # coordinates agree to ~1e-6; on 0..255 noise images a 1e-4 pixel shift moves values by ~1e-2
  assert float((out - ref).abs().max()) < 0.05
  assert float((out - ref).abs().mean()) < 2e-3


@pytest.mark.gpu
def test_apply_tps_produces_the_input_contract():
  from imm_b200.utils.tps_sampler import TPSRandomSampler, apply_tps
  B, R = 2, 128
  img = O.synthetic_inputs(B, R)['image']
  mask = O.smooth_mask(R, R).view(1, R, R, 1).repeat(B, 1, 1, 1)
  rng = np.random.RandomState(9)
  wt = torch.from_numpy(np.stack([T.sample_tps_w(10, 10, (0.001, 0.005), 0.0, 0.0, 0.1, rng) for _ in range(B)]).astype(np.float32))
  ws = torch.from_numpy(np.stack([T.sample_tps_w(10, 10, (0.001, 0.01), 0.0, 0.0, 0.1, rng) for _ in range(B)]).astype(np.float32))
  ref = T.apply_tps(img, mask, wt, ws)

  class Fixed(TPSRandomSampler):
    def __init__(self, w):
      TPSRandomSampler.__init__(self, R, R, pad=False)
      self._w = w

    def _get_params(self, batch_size):
      return self._w.cuda()
  out = apply_tps(img.cuda(), mask.cuda(), Fixed(wt), Fixed(ws))
  for k in ('image', 'future_image', 'mask'):
    assert out[k].shape == ref[k].shape
    assert float((out[k].cpu() - ref[k]).abs().max()) < 0.05, k
  assert out['mask'].shape == (B, R, R, 1) and float(out['mask'].max()) <= 1.0 + 1e-5


@pytest.mark.gpu
def test_tps_dataset_feeds_the_model():
  """The GPU pair generator output drives one training step through the public API."""
  from imm_b200.datasets.tps_dataset import TPSDataset
  from imm_b200.models.imm_model import IMMModel
  from imm_b200.utils.box import default_model_config
  from imm_b200.utils.synthetic import synthetic_vgg_caffe_dict
  ds = TPSDataset(seed=1).get_dataset(2)
  batch = ds()
  assert batch['image'].shape == (2, 128, 128, 3) and batch['mask'].shape == (2, 128, 128, 1)
  assert float((batch['image'] - batch['future_image']).abs().mean()) > 0.1       # the two views differ
  model = IMMModel(default_model_config(10), vgg_data=synthetic_vgg_caffe_dict(1))
  _, loss, _ = model.build(batch, True)
  model.engine.backward()
  model.engine.optimizer_step()
  assert torch.isfinite(loss).item()
