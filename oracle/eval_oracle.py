"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Restatement of the landmark-regression metric of scripts/test.py:36-63
(ridge regression with alpha = 0 and no intercept == ordinary least squares; inter-ocular normalised error),
written with numpy's lstsq so that it is independent of sklearn.  PARITY UNPINNED (the reference has no tests)."""
import numpy as np


def regression_error(yx_train, gt_train, yx_test, gt_test, im_size):
  def conv(yx):
    return (((yx + 1) / 2.0) * np.array(im_size)).reshape(yx.shape[0], -1)
  X_tr, X_te = conv(yx_train), conv(yx_test)
  Y_tr = gt_train.astype(np.float32).reshape(gt_train.shape[0], -1)
  W, _, _, _ = np.linalg.lstsq(X_tr.astype(np.float64), Y_tr.astype(np.float64), rcond=None)
  pred = (X_te.astype(np.float64) @ W).reshape(gt_test.shape)
  gt = gt_test.astype(np.float32)
  iod = np.sqrt(((gt[:, 0, :] - gt[:, 1, :]) ** 2).sum(-1))
  dist = np.sqrt(((gt - pred) ** 2).sum(-1))
  return float(np.mean(dist / iod[:, None]))
