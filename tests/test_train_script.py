"""End-to-end through the reference-facing CLI: scripts/train.py with a reference-schema YAML, checkpoint layout
(TF variable names) and the restore semantics of cnn_train_multi.py:404-433."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CFG = '''
name: t
logdir: %s
vgg16_path: /nonexistent/vgg16.caffemodel.h5
training:
  ncheckpoint: 2
  gradclip: 1.0
  dset: synthetic
  logdir: ${logdir}/${name}
  datadir: none
  batch: 4
  allow_growth: True
  optim: Adam
  lr: {start_val: 0.001, step: 100000, decay: 0.95}
model:
  gauss_std: 0.10
  gauss_mode: 'rot'
  n_maps: 10
  n_filters: 32
  n_filters_render: 32
  renderer_stride: 2
  min_res: 16
  reconstruction_loss: perceptual
  perceptual:
    l2: True
    comp: ['input', 'conv1_2', 'conv2_2', 'conv3_2', 'conv4_2', 'conv5_2']
    net_file: ${vgg16_path}
  loss_mask: True
  channels_bug_fix: True
'''


def test_train_script_runs_checkpoints_and_restores(tmp_path):
  cfg = tmp_path / 'exp.yaml'
  cfg.write_text(CFG % str(tmp_path))
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'train.py'), '--configs', str(cfg), '--num-steps', '3', '--synthetic'],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=ROOT, timeout=900)
  assert out.returncode == 0, out.stdout[-3000:]
  assert 'step -1, loss' in out.stdout and 'step 2, loss' in out.stdout       # global_step starts at -1 (train.py:87-89)
  assert 'Avg. samples per second' in out.stdout
  assert any(f.startswith('events.out.tfevents') for f in os.listdir(str(tmp_path / 't')))      # train summaries
  # saver.save(session, <logdir>/model.ckpt, global_step=step): TensorBundle files + the `checkpoint` state file
  from imm_b200.utils import tf_checkpoint
  ck = tmp_path / 't' / 'model.ckpt-2'
  assert (tmp_path / 't' / 'model.ckpt-2.index').exists() and (tmp_path / 't' / 'model.ckpt-2.data-00000-of-00001').exists()
  assert tf_checkpoint.latest_checkpoint(str(tmp_path / 't')) == str(ck)
  reader = tf_checkpoint.CheckpointReader(str(ck))
  sd = {k: torch.from_numpy(v) for k, v in reader.read_all().items()}
  assert all(v.dtype == torch.float32 for v in sd.values()) and reader.get_variable_to_shape_map()['global_step'] == []
  for k in ('model/image_encoder/encoder/conv_1/conv_1/w', 'model/pose_encoder/conv_1/conv_1/b',
            'model/renderer/conv_7/batch_normalization/moving_variance', 'SelfSupReconstructionLoss/conv3_2_agg',
            'SelfSupReconstructionLoss/vgg16/conv5_2/weights', 'global_step', 'beta1_power',
            'model/renderer/conv_1/conv_1/w/Adam_1'):
    assert k in sd, k
  assert tuple(sd['model/renderer/conv_1/conv_1/w'].shape) == (3, 3, 266, 256)      # HWIO on disk (base_model.py:110)
  # restore into a fresh model: 'model' = MODEL_VARIABLES only (BN variables keep their initial values)
  from imm_b200.models.imm_model import IMMModel
  from imm_b200.utils.box import default_model_config
  from imm_b200.utils.synthetic import synthetic_inputs, synthetic_vgg_caffe_dict
  m = IMMModel(default_model_config(10), vgg_data=synthetic_vgg_caffe_dict(1), seed=123)
  m.build(synthetic_inputs(4), False)
  m.load_state_dict(sd, vars_to_restore='model')
  k = 'model/renderer/conv_3/conv_3/w'
  assert torch.equal(m.engine.params[k].cpu(), sd[k])
  g = 'model/renderer/conv_3/batch_normalization/gamma'
  assert not torch.equal(m.engine.params[g].cpu(), sd[g]) or float(sd[g].std()) == 0.0
  m.load_state_dict(sd, vars_to_restore='all')
  assert torch.equal(m.engine.params[g].cpu(), sd[g])
  assert torch.equal(m.engine.adam_v[k].cpu(), sd[k + '/Adam_1'])
  # a full restore also takes the frozen tower from the checkpoint (its weights are global variables too): the fresh
  # model above was built with VGG seed 1 like the training run, so perturb it first to see the restore act
  vk = 'SelfSupReconstructionLoss/vgg16/conv3_2/weights'
  m.engine.vgg_params[vk].mul_(0.5)
  m.load_state_dict(sd, vars_to_restore='all')
  assert torch.equal(m.engine.vgg_params[vk].cpu(), sd[vk])
  assert m.engine.global_step == float(sd['global_step']) == 3.0   # -1 + 4 applied steps; the FILE is named by the loop step (2)
  assert m.engine.adam_t == 4          # steps -1..2 applied; recovered from beta1_power = 0.9^(t+1)
  # --checkpoint restore through the CLI path (cnn_train_multi.py:404-433) continues from global_step + 0
  out = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'train.py'), '--configs', str(cfg), '--num-steps', '4', '--synthetic',
                        '--checkpoint', str(ck), '--restore-optim'],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, cwd=ROOT, timeout=900)
  assert out.returncode == 0, out.stdout[-3000:]
  assert 'RESTORING MODEL from' in out.stdout and 'step 3, loss' in out.stdout and 'step 1, loss' not in out.stdout
