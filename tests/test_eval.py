"""evaluation loop + landmark-regression metric."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import eval_oracle as EO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_test_script():
  spec = importlib.util.spec_from_file_location('imm_test_script', os.path.join(ROOT, 'scripts', 'test.py'))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


def test_regression_error_known_answers():
  T = _load_test_script()
  rng = np.random.RandomState(0)
  K, n = 6, 200
  yx = rng.rand(n, K, 2) * 2 - 1
  X = (((yx + 1) / 2.0) * np.array([128, 128])).reshape(n, -1)
  A = rng.randn(2 * K, 10)
  gt = (X @ A).reshape(n, 5, 2)                              # exactly linear in the landmarks -> error 0
  tr = {'gauss_yx': yx[:150], 'future_landmarks': gt[:150]}
  te = {'gauss_yx': yx[150:], 'future_landmarks': gt[150:]}
  assert T.regression_error(tr, te, [128, 128]) < 1e-4
  # shift landmark 2 of the test set by (3, 4) pixels: its distance is 5 px; mean over 5 landmarks / iod
  te2 = {'gauss_yx': yx[150:], 'future_landmarks': gt[150:].copy()}
  te2['future_landmarks'][:, 2, :] += np.array([3.0, 4.0])
  g = te2['future_landmarks'].astype(np.float32)
  iod = np.sqrt(((g[:, 0] - g[:, 1]) ** 2).sum(-1))
  expect = np.mean(5.0 / iod) / 5.0
  got = T.regression_error(tr, te2, [128, 128])
  np.testing.assert_allclose(got, expect, rtol=1e-3)
  np.testing.assert_allclose(got, EO.regression_error(yx[:150], gt[:150], yx[150:], te2['future_landmarks'], [128, 128]), rtol=1e-3)


class _BlobDataset(object):
  """Finite labelled toy set: images with 5 bright blobs at annotated positions."""
  image_size = [128, 128]

  def __init__(self, n_batches, seed):
    self.n_batches, self.seed = n_batches, seed

  def get_dataset(self, batch_size, repeat=False, shuffle=False, num_preprocess_threads=12):
    rng = np.random.RandomState(self.seed)
    state = {'i': 0}
    yy, xx = np.mgrid[0:128, 0:128].astype(np.float32)

    def nxt():
      if state['i'] >= self.n_batches:
        return None
      state['i'] += 1
      lm = rng.rand(batch_size, 5, 2).astype(np.float32) * 80 + 24
      img = np.zeros((batch_size, 128, 128, 3), np.float32)
      for b in range(batch_size):
        for k in range(5):
          img[b] += (255.0 * np.exp(-((yy - lm[b, k, 0]) ** 2 + (xx - lm[b, k, 1]) ** 2) / 50.0))[..., None]
      img = np.clip(img, 0, 255)
      t = torch.from_numpy(img)
      return {'image': t, 'future_image': t, 'mask': torch.ones(batch_size, 128, 128, 1), 'future_landmarks': lm}
    return nxt


@pytest.mark.gpu
def test_eval_loop_restores_checkpoint_and_collects_tensors(tmp_path):
  from imm_b200.models.imm_model import IMMModel
  from imm_b200.utils.box import Box, default_model_config
  from imm_b200.utils.synthetic import synthetic_vgg_caffe_dict
  T = _load_test_script()
  cfg = default_model_config(10)
  m = IMMModel(cfg, vgg_data=synthetic_vgg_caffe_dict(1), seed=7)
  b = _BlobDataset(1, 0).get_dataset(4)()
  m.build(b, True)
  m.engine.backward()
  m.engine.optimizer_step()
  torch.save(m.state_dict(), str(tmp_path / 'model.ckpt-0'))
  ref_w = m.engine.params['model/pose_encoder/conv_1/conv_1/w'].cpu().clone()
  err = T.evaluate(IMMModel, 'model.ckpt-0', cfg, Box({'logdir': str(tmp_path), 'allow_growth': True}),
                   _BlobDataset(3, 1), _BlobDataset(2, 2), batch_size=4,
                   net_kwargs={'vgg_data': synthetic_vgg_caffe_dict(1), 'seed': 99})
  assert np.isfinite(err) and err > 0
  # the loop itself: tensors are lists of per-batch arrays; eval mode leaves the restored weights untouched
  from imm_b200.eval import eval_imm
  res = eval_imm.evaluate(_BlobDataset(2, 3), IMMModel, cfg, 'model.ckpt-0', Box({'logdir': str(tmp_path)}), batch_size=4,
                          eval_tensors=['gauss_yx', 'future_landmarks', 'heatmaps'],
                          net_kwargs={'vgg_data': synthetic_vgg_caffe_dict(1), 'seed': 5})
  assert len(res['gauss_yx']) == 2 and res['gauss_yx'][0].shape == (4, 10, 2) and res['heatmaps'][0].shape == (4, 16, 16, 10)
  assert np.all(np.abs(res['gauss_yx'][0]) <= 1.0)
