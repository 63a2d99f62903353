"""Landmark-regression evaluation -- mirror of the reference's scripts/test.py.

K unsupervised landmarks -> 5 annotated landmarks by ridge regression without bias (sklearn, alpha=0, as the
reference), error = mean point distance normalised by the inter-ocular distance (scripts/test.py:36-63)."""
from __future__ import print_function

import os.path as osp
import sys

import numpy as np
import sklearn.linear_model

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))

from imm_b200.eval import eval_imm  # noqa: E402


def regression_error(train_tensors, test_tensors, im_size, bias=False):
  """scripts/test.py:36-63 on already collected tensors (dicts with 'gauss_yx' [N,K,2] in [-1,1] and
  'future_landmarks' [N,5,2] in pixels)."""
  def convert_landmarks(tensors, im_size):
    landmarks = tensors['gauss_yx']
    landmarks_gt = tensors['future_landmarks'].astype(np.float32)
    im_size = np.array(im_size)
    landmarks = ((landmarks + 1) / 2.0) * im_size
    n_samples = landmarks.shape[0]
    return landmarks.reshape((n_samples, -1)), landmarks_gt.reshape((n_samples, -1))

  X_train, y_train = convert_landmarks(train_tensors, im_size)
  X_test, y_test = convert_landmarks(test_tensors, im_size)
  regr = sklearn.linear_model.Ridge(alpha=0.0, fit_intercept=bias)
  _ = regr.fit(X_train, y_train)
  y_predict = regr.predict(X_test)
  landmarks_gt = test_tensors['future_landmarks'].astype(np.float32)
  landmarks_regressed = y_predict.reshape(landmarks_gt.shape)
  eyes = landmarks_gt[:, :2, :]
  occular_distances = np.sqrt(np.sum((eyes[:, 0, :] - eyes[:, 1, :]) ** 2, axis=-1))
  distances = np.sqrt(np.sum((landmarks_gt - landmarks_regressed) ** 2, axis=-1))
  return np.mean(distances / occular_distances[:, None])


def evaluate(net, net_file, model_config, training_config, train_dset, test_dset, batch_size=100, bias=False,
             net_kwargs=None):
  """scripts/test.py:18-65."""
  def run(dset):
    results = eval_imm.evaluate(dset, net, model_config, net_file, training_config, batch_size=batch_size,
                                random_seed=0, eval_tensors=['gauss_yx', 'future_landmarks'], net_kwargs=net_kwargs)
    return {k: np.concatenate(v) for k, v in results.items()}
  train_tensors = run(train_dset)
  test_tensors = run(test_dset)
  return regression_error(train_tensors, test_tensors, train_dset.image_size, bias=bias)


if __name__ == '__main__':
  raise SystemExit('MAFL / AFLW are not reachable offline: call evaluate() with dataset objects that provide '
                   "get_dataset() and 'future_landmarks' (see tests/test_eval.py for a synthetic example)")
