"""Regenerates tests/golden/c1_fp64.npz: the fp64 oracle on BASELINE config 1 (CelebA-10pts, batch 2,
128x128 synthetic pairs, one fwd+loss+bwd+clip+Adam step, seed 0).  The reference itself cannot run here
(TensorFlow 1.10 is not installable), so this fixture pins the ORACLE against regressions; the oracle's
TF semantics are pinned separately by tests/test_oracle_kat.py.   Usage: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import imm_oracle as O  # noqa: E402


def run(dtype=torch.float64, batch=2, n_maps=10, seed=0):
  st = O.init_state(O.State(n_maps=n_maps), seed=seed).clone(dtype)
  inp = {k: v.to(dtype) for k, v in O.synthetic_inputs(batch, seed=seed).items()}
  r = O.train_step(st, inp)
  out = r['out']
  d = {'loss': r['loss'].detach().numpy(), 'lr': np.float64(r['lr']),
       'level_losses': np.array([float(x) for x in out['level_losses']]),
       'gauss_yx': out['gauss_yx'].detach().numpy(),
       'pred_sub': out['future_im_pred'].detach().numpy()[:, ::8, ::8, :],
       'pred_sum': out['future_im_pred'].detach().sum().numpy(),
       'pred_sqsum': (out['future_im_pred'].detach() ** 2).sum().numpy(),
       'heatmaps_sub': out['heatmaps'].detach().numpy()[:, ::4, ::4, :]}
  names = list(st.params.keys())
  d['param_names'] = np.array(names)
  d['grad_norms'] = np.array([float(r['grads'][k].norm()) for k in names])
  d['param_norms_after'] = np.array([float(st.params[k].norm()) for k in names])
  d['agg_after'] = np.array([float(st.buffers['SelfSupReconstructionLoss/%s_agg' % n]) for n in st.perceptual_comp])
  d['mv_after_enc1'] = st.buffers['model/image_encoder/encoder/conv_1/batch_normalization/moving_variance'].numpy()
  return d


if __name__ == '__main__':
  d = run()
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'c1_fp64.npz')
  np.savez_compressed(path, **d)
  print('wrote', path, os.path.getsize(path), 'bytes; loss', float(d['loss']))
