"""CLI entry -- mirror of the reference's scripts/train.py (same flags; same YAML configs).

  python scripts/train.py --configs configs/paths/default.yaml configs/experiments/celeba-10pts.yaml [--ngpus N]

Multi-GPU: launch with torchrun (one process per GPU); --ngpus is then the world size.  `training.dset` selects the
dataset class (celeba | aflw: host readers + GPU TPS warps; synthetic | tps: seeded streams); a missing dataset
directory or VGG16 weight file is an error unless --synthetic asks for the seeded stand-ins explicitly."""
from __future__ import print_function

import argparse
import os
import os.path as osp
import sys

sys.path.insert(0, osp.dirname(osp.dirname(osp.abspath(__file__))))

import torch  # noqa: E402

import imm_b200.train.cnn_train_multi as tru  # noqa: E402
from imm_b200.utils.dataset_import import import_dataset  # noqa: E402
from imm_b200.models.imm_model import IMMModel  # noqa: E402
from imm_b200.utils.box import read_configs  # noqa: E402
from imm_b200.utils.synthetic import synthetic_vgg_caffe_dict  # noqa: E402


class model_factory():
  """Factory which can be used to instantiate models (scripts/train.py:30-40)."""

  def __init__(self, network, **kwargs):
    self.network = network
    self.net_args = kwargs

  def create(self):
    return self.network(**self.net_args)


def load_configs(file_names):
  return read_configs(file_names)


def main(args):
  config = load_configs(args.configs)
  train_config = config.training
  rank, local_rank, world = tru.init_distributed()
  ngpus = max(args.ngpus, world)
  gpus = list(range(ngpus))
  log_dir = train_config.logdir
  NUM_STEPS = args.num_steps
  GRAD_CLIP = train_config.gradclip
  checkpoint_fname = args.checkpoint if args.checkpoint is not None else osp.join(log_dir, 'INVALID')
  opts = {'gpu_ids': gpus, 'log_dir': log_dir, 'n_summary': 10,
          'n_test': train_config.n_test if hasattr(train_config, 'n_test') else 500,
          'n_checkpoint': train_config.ncheckpoint, 'batch_size': train_config.batch}
  batch_size = train_config.batch
  assert batch_size % world == 0
  lr = tru.exponential_decay(train_config.lr.start_val, train_config.lr.step, train_config.lr.decay,
                             staircase=True, lr_multiple=args.lr_multiple)
  if train_config.optim.lower() != 'adam':
    raise ValueError('Optimizer = %s not suppoerted' % train_config.optim)
  optim = tru.AdamOptimizer(lr, name='Adam')
  vgg = None
  if args.synthetic:
    if rank == 0:
      print('--synthetic: seeded synthetic VGG16 weights and the synthetic image-pair stream')
    vgg = synthetic_vgg_caffe_dict(1)
  elif not osp.exists(str(config.model.perceptual.net_file)) and args.checkpoint is None:
    raise IOError('VGG16 weights %s not found (build_vgg16.py:16); pass --synthetic for seeded stand-in weights, or '
                  '--checkpoint --restore-optim to take the frozen tower from a full checkpoint'
                  % config.model.perceptual.net_file)
  factory = model_factory(IMMModel, config=config.model, global_step=args.reset_global_step,
                          device='cuda:%d' % local_rank, world_size=world, vgg_data=vgg)
  # dataset class by name (scripts/train.py:116-147), default / configured constructor parameters and subsets
  dset_name = 'synthetic' if args.synthetic else train_config.dset
  dset_class = import_dataset(dset_name)
  train_params, test_params = {}, {}
  train_subset, test_subset = 'train', 'test'
  if hasattr(train_config, 'train_dset_params'):
    train_params.update(train_config.train_dset_params)
    train_subset = train_params.pop('subset', train_subset)
  if hasattr(train_config, 'test_dset_params'):
    test_params.update(train_config.test_dset_params)
    test_subset = test_params.pop('subset', test_subset)
  if hasattr(train_config, 'max_test_samples'):
    raise ValueError('max_test_samples attribute deprecated')
  if dset_name in ('celeba', 'aflw'):
    if not osp.isdir(str(train_config.datadir)):
      raise IOError('dataset directory %s not found (training.dset = %s); pass --synthetic to train on the seeded '
                    'synthetic pair stream instead' % (train_config.datadir, dset_name))
    train_params['device'] = test_params['device'] = 'cuda:%d' % local_rank
  dset = dset_class(train_config.datadir, subset=train_subset, **train_params).get_dataset(
      batch_size // world, repeat=True, shuffle=False, num_preprocess_threads=12, rank=rank)
  test_dset = None
  if dset_name in ('celeba', 'aflw'):
    test_obj = dset_class(train_config.datadir, subset=test_subset, **test_params)
    test_dset = lambda: test_obj.get_dataset(batch_size // world, repeat=False, shuffle=False, num_preprocess_threads=12)
  loss, train_op, _, _, model = tru.setup_training(opts, None, optim, dset, True, factory, args.reset_global_step,
                                                   clip_value=GRAD_CLIP, split_gpus=False)
  model.build(dset(), False, build_loss=False) if vgg is None else model.build(dset(), False)   # instantiate the engine
  if vgg is None and osp.exists(str(config.model.perceptual.net_file)):
    model.load_vgg()
  restore_vars = 'all' if args.restore_optim else 'model'
  tru.train_loop(opts, None, loss, dset, True, None, train_op, None, None, NUM_STEPS, args.reset_global_step,
                 checkpoint_fname, test_dataset=test_dset, ignore_missing_vars=args.ignore_missing_vars,
                 reset_global_step=args.reset_global_step, vars_to_restore=restore_vars, exclude_vars=[],
                 allow_growth=getattr(train_config, 'allow_growth', True), model=model)


if __name__ == '__main__':
  parser = argparse.ArgumentParser(description='Train Unsupervised Sequence Model')
  parser.add_argument('--configs', nargs='+', default=[], help='Paths to the config files.')
  parser.add_argument('--ngpus', type=int, default=1, required=False, help='Number of GPUs to use for training.')
  parser.add_argument('--lr-multiple', type=float, default=1, help='multiplier on the learning rate.')
  parser.add_argument('--checkpoint', type=str, default=None, help='checkpoint file-name of the *FULL* model to restore.')
  parser.add_argument('--restore-optim', action='store_true', help='Restore the optimizer variables.')
  parser.add_argument('--reset-global-step', type=int, default=-1, help='Force the value of global step.')
  parser.add_argument('--ignore-missing-vars', action='store_true', help='Skip re-storing vars not in the checkpoint file.')
  parser.add_argument('--num-steps', type=int, default=30000000, help='(extension) stop after this many steps.')
  parser.add_argument('--synthetic', action='store_true',
                      help='(extension) seeded synthetic image pairs and synthetic VGG16 weights instead of training.dset / perceptual.net_file.')
  main(parser.parse_args())
