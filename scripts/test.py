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
  """The landmark-regression metric of the reference (scripts/test.py:36-63) on already collected tensors: dicts with
  'gauss_yx' [N,K,2] in [-1,1] (y,x) and 'future_landmarks' [N,5,2] in pixels.  Unsupervised landmarks are mapped to
  pixel units, a ridge regressor with alpha = 0 (ordinary least squares; no intercept unless `bias`) is fitted on the
  training set, and the test error is the point distance normalised by the inter-ocular distance (annotated landmarks 0
  and 1 are the eyes), averaged over landmarks and samples."""
  size = np.asarray(im_size, dtype=np.float64)

  def features(t):
    yx = np.asarray(t['gauss_yx'], dtype=np.float64)
    return (0.5 * (yx + 1.0) * size).reshape(len(yx), -1)

  gt_train = np.asarray(train_tensors['future_landmarks'], dtype=np.float32)
  gt_test = np.asarray(test_tensors['future_landmarks'], dtype=np.float32)
  model = sklearn.linear_model.Ridge(alpha=0.0, fit_intercept=bias)
  model.fit(features(train_tensors), gt_train.reshape(len(gt_train), -1))
  regressed = model.predict(features(test_tensors)).reshape(gt_test.shape)
  inter_ocular = np.linalg.norm(gt_test[:, 0] - gt_test[:, 1], axis=-1)
  per_point = np.linalg.norm(gt_test - regressed, axis=-1)
  return float(np.mean(per_point / inter_ocular[:, None]))


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


def main(args):
  """scripts/test.py:68-150: builds the regressor-train / test datasets, finds the checkpoint and prints the error."""
  from imm_b200.models.imm_model import IMMModel
  from imm_b200.utils.box import read_configs
  from imm_b200.utils.dataset_import import import_dataset
  config = read_configs([args.paths_config, osp.join('configs', 'experiments', args.experiment_name + '.yaml')])
  im_size = [args.im_size, args.im_size]

  def make(which, subset):
    if which == 'mafl':
      return import_dataset('celeba')(config.training.datadir, dataset='mafl', subset=subset, order_stream=True,
                                      tps=False, image_size=im_size)
    if which == 'aflw':
      return import_dataset('aflw')(config.training.datadir, subset=subset, order_stream=True, tps=False,
                                    image_size=im_size)
    raise ValueError('Dataset %s not supported.' % which)
  train_dset = make(args.train_dataset, 'train')
  test_dset = make(args.test_dataset, args.test_split)
  net_file = 'model.ckpt' if args.iteration is None else 'model.ckpt-' + str(args.iteration)
  checkpoint_file = osp.join(config.training.logdir, net_file + '.index')
  if not osp.isfile(checkpoint_file):
    raise ValueError('Checkpoint file %s not found.' % checkpoint_file)
  mean_error = evaluate(IMMModel, net_file, config.model, config.training, train_dset, test_dset,
                        batch_size=args.batch_size, bias=args.bias)
  params = config.training.train_dset_params
  model_dataset = params.dataset if hasattr(params, 'dataset') else config.training.dset
  print('')
  print('========================= RESULTS =========================')
  print('model trained in unsupervised way on %s dataset' % model_dataset)
  print('regressor trained on %s training set' % args.train_dataset)
  print('error on %s datset %s set: %.5f (%.3f percent)' % (args.test_dataset, args.test_split, mean_error,
                                                            mean_error * 100.0))
  print('===========================================================')


if __name__ == '__main__':
  import argparse
  parser = argparse.ArgumentParser(description='Test model on face datasets.')
  parser.add_argument('--experiment-name', type=str, required=True, help='Name of the experiment to evaluate.')
  parser.add_argument('--train-dataset', type=str, required=True, help='Training dataset for regressor (mafl|aflw).')
  parser.add_argument('--test-dataset', type=str, required=True, help='Testing dataset for regressed landmarks (mafl|aflw).')
  parser.add_argument('--paths-config', type=str, default='configs/paths/default.yaml', help='Path to the paths config.')
  parser.add_argument('--iteration', type=int, default=None, help='Checkpoint iteration to evaluate.')
  parser.add_argument('--test-split', type=str, default='test', help='Test split (val|test).')
  parser.add_argument('--buffer-name', type=str, default=None, help='(accepted for CLI compatibility; unused)')
  parser.add_argument('--im-size', type=int, default=128, help='Image size.')
  parser.add_argument('--bias', action='store_true', help='Use bias in the regressor.')
  parser.add_argument('--batch-size', type=int, default=100, help='batch_size')
  main(parser.parse_args())
