"""Evaluation loop -- host-side mirror of imm/eval/eval_imm.py:evaluate.

Restores a checkpoint (tensors keyed by the reference's TF variable names), runs the inference-mode forward
(training_pl=False: BN uses the moving statistics, no state is updated) over a finite dataset and collects the
requested tensors as lists of numpy arrays, exactly the structure scripts/test.py consumes."""
import os
import time
from datetime import datetime

import numpy as np
import torch


def evaluate(dataset_instance, net, net_config, net_file, training_opts, batch_size=100, random_seed=0,
             eval_tensors=None, eval_loss=False, eval_summaries=False, eval_metrics=False, net_kwargs=None):
  """eval_imm.py:25-143.  dataset_instance.get_dataset(batch_size, repeat=False, ...) must return a callable that
  yields input dicts and returns None (or raises StopIteration) when the set is exhausted."""
  np.random.seed(random_seed)
  test_dataset = dataset_instance.get_dataset(batch_size, repeat=False, shuffle=False, num_preprocess_threads=12)
  net_instance = net(net_config, **(net_kwargs or {}))
  ckpt = os.path.join(training_opts.logdir, net_file) if not os.path.isabs(net_file) else net_file
  tensors_results, restored, test_iter = {}, False, 0
  while True:
    try:
      inputs = test_dataset()
    except StopIteration:
      inputs = None
    if inputs is None:
      print('iteration through test set finished')
      break
    start_time = time.time()
    if not restored:
      net_instance.build(inputs, False, output_tensors=True, build_loss=False)      # instantiates the engine
      if not (os.path.exists(ckpt) or os.path.exists(ckpt + '.index')):               # eval_imm.py:81
        raise Exception('model file does not exist at: ' + ckpt)                     # eval_imm.py:94-95
      print('RESTORING MODEL from: ' + ckpt)
      # every global variable present in the checkpoint is restored (eval_imm.py:81-94)
      net_instance.restore_checkpoint(ckpt, vars_to_restore='all', ignore_missing_vars=True)
      restored = True
    _, loss, _, tensors = net_instance.build(inputs, False, output_tensors=True, build_loss=eval_loss)
    tensors.update(net_instance.get_collection('tensors'))
    names = list(tensors.keys()) if eval_tensors is None else list(eval_tensors)
    for name in names:
      v = tensors[name]
      v = v() if callable(v) else v
      v = v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)
      tensors_results.setdefault(name, []).append(v)
    loss_value = float(loss.item()) if eval_loss else 0
    duration = time.time() - start_time
    print('test: %s: step %d, loss = %.4f (%.1f examples/sec) %.3f sec/batch'
          % (datetime.now(), int(net_instance.engine.global_step), loss_value, batch_size / float(duration), duration))
    test_iter += 1
  return tensors_results
