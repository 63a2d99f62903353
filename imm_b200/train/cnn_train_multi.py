"""Training runtime -- host-side mirror of imm/train/cnn_train_multi.py.

Reference: in-graph tower replication, gradients averaged on the CPU (cnn_train_multi.py:66-106,109-192).
Here: one process per GPU (torchrun), one NCCL all-reduce of the flat gradient buffer over NVLink per step,
then the fused per-tensor clip_by_norm + TF-Adam kernel.  Function names / argument meaning follow the
reference; `graph` is accepted and ignored.  file:line citations are under /root/reference."""
import os
import time
from datetime import datetime

import numpy as np
import torch
import torch.distributed as dist

from ..utils import tf_checkpoint


class AdamOptimizer(object):
  """tf.train.AdamOptimizer(lr, name='Adam') stand-in (scripts/train.py:98): holds the hyper-parameters; the
  update itself is immb_adam_apply.  `learning_rate` may be a float or a callable(global_step) -> float."""

  def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8, name='Adam'):
    self.learning_rate, self.beta1, self.beta2, self.epsilon, self.name = learning_rate, beta1, beta2, epsilon, name

  def lr(self, global_step):
    return self.learning_rate(global_step) if callable(self.learning_rate) else float(self.learning_rate)


def exponential_decay(start_val, decay_steps, decay_rate, staircase=True, lr_multiple=1.0):
  """lr_multiple * tf.train.exponential_decay(...) (scripts/train.py:92-96)."""
  def fn(global_step):
    p = global_step / float(decay_steps)
    if staircase:
      p = np.floor(p)
    return lr_multiple * start_val * decay_rate ** p
  return fn


def world_info():
  ws = int(os.environ.get('WORLD_SIZE', '1'))
  return int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0')), ws


def init_distributed(backend='nccl'):
  rank, local_rank, ws = world_info()
  if ws > 1 and not dist.is_initialized():
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    if backend == 'nccl':
      torch.cuda.set_device(local_rank)
    dist.init_process_group(backend=backend, rank=rank, world_size=ws)
  return rank, local_rank, ws


def average_gradients(flat_grads, clip_value=None):
  """cnn_train_multi.py:66-106: mean over towers (the synchronisation point).  One all-reduce(sum) on the flat
  bucket; the 1/N scale and the per-tensor clip are fused into the optimiser kernels (gscale, clip)."""
  if dist.is_initialized() and dist.get_world_size() > 1:
    dist.all_reduce(flat_grads, op=dist.ReduceOp.SUM)
  return flat_grads


def tower_loss(inputs, training_pl, model, scope=None):
  """cnn_train_multi.py:37-64."""
  _, loss, avg_ops = model.build(inputs, training_pl, costs_collection='costs', scope=scope, var_device='/cpu:0')
  return loss


class _InputPrefetcher(object):
  """One-deep input prefetch, the role `dataset.prefetch` + the feedable iterator play in the reference
  (tps_dataset.py:98-131, cnn_train_multi.py:445-450): the NEXT batch's host->device copies are enqueued on a copy
  stream while the current step's kernels run, so H2D never sits on the critical path.  Host tensors should be pinned
  (pageable memory makes the copy synchronous, which is still correct).  CUDA tensors pass through untouched."""

  def __init__(self, inputs_fn, device):
    self.inputs_fn, self.device = inputs_fn, torch.device(device)
    self.stream = torch.cuda.Stream(device=self.device)
    self.staged = None

  def _stage(self):
    inputs = self.inputs_fn()
    if inputs is None:
      self.staged = None
      return
    out = {}
    with torch.cuda.stream(self.stream):
      for k, v in inputs.items():
        out[k] = v.to(self.device, non_blocking=True) if isinstance(v, torch.Tensor) and not v.is_cuda else v
    ev = torch.cuda.Event()
    ev.record(self.stream)
    self.staged = (out, ev)

  def next(self):
    if self.staged is None:
      self._stage()
    if self.staged is None:
      return None
    out, ev = self.staged
    cur = torch.cuda.current_stream(self.device)
    cur.wait_event(ev)
    for v in out.values():
      if isinstance(v, torch.Tensor) and v.is_cuda:
        v.record_stream(cur)          # allocated on the copy stream, consumed on the compute stream
    self.staged = None
    return out

  def prefetch(self):
    if self.staged is None:
      self._stage()


def _make_train_op(model, optim, inputs_fn, clip_value):
  state = {}

  def train_op():
    pf = state.get('pf')
    if pf is None:
      pf = state['pf'] = _InputPrefetcher(inputs_fn, model._device)
    inputs = pf.next()
    gs = model.engine.global_step if model.engine is not None else float(model._global_step if model._global_step is not None else -1)
    world = model.engine.world_size if model.engine is not None else model._world
    loss = model.train_step(inputs, clip_value, lr=optim.lr(gs), beta1=optim.beta1, beta2=optim.beta2, eps=optim.epsilon,
                            allreduce=average_gradients if world > 1 else None)
    pf.prefetch()                  # next batch's H2D overlaps this step's kernels (the host is ahead of the GPU here)
    return loss        # cost EMAs (model._avg_ops) are evaluated lazily by the logger: they need a D2H read
  return train_op


def train_single(opts, graph, optim, inputs, training_pl, model_factory, global_step, clip_value=None):
  """cnn_train_multi.py:195-250."""
  model = model_factory.create()
  train_op = _make_train_op(model, optim, inputs, clip_value)
  loss = lambda: model.engine.total_loss
  return loss, train_op, None, None, model


def train_multi(opts, graph, optim, inputs, training_pl, model_factory, global_step, clip_value=None):
  """cnn_train_multi.py:109-192: batch split across GPUs (:132), one tower per GPU, gradient mean, clip, apply.
  `inputs` yields this rank's slice of the global batch (utils.split_tensors semantics, imm/utils/utils.py:113).
  BN statistics stay per-replica (no SyncBN), as in the reference (:155,166)."""
  num_gpus = len(opts['gpu_ids'])
  assert opts['batch_size'] % num_gpus == 0, ('Batch size must be divisible by number of GPUs')
  return train_single(opts, graph, optim, inputs, training_pl, model_factory, global_step, clip_value)


def setup_training(opts, graph, optim, inputs, training_pl, model_factory, global_step, clip_value=None,
                   split_gpus=False):
  """cnn_train_multi.py:342-368."""
  if split_gpus:
    raise NotImplementedError('split_gpus component placement is not enabled by any shipped config')
  num_gpus = len(opts['gpu_ids'])
  if num_gpus == 0:
    raise RuntimeError('training on CPU is not available: the CUDA path has no CPU fallback')
  if num_gpus == 1:
    return train_single(opts, graph, optim, inputs, training_pl, model_factory, global_step, clip_value)
  return train_multi(opts, graph, optim, inputs, training_pl, model_factory, global_step, clip_value)


def run_test_pass(opts, model, test_dataset, step, summary=None):
  """The periodic test pass of train_loop (cnn_train_multi.py:472-506): one full pass over the (finite) test set with
  training_pl = False -- BN uses the moving statistics, no state is updated -- logging the loss of every batch; the
  streaming `loss` metric (imm_model.py:392, op_utils.py:80-91: mean over the pass) goes to the summary writer.
  test_dataset: a callable yielding input dicts and None when exhausted, or a zero-argument factory returning one (the
  re-initialisable iterator of the reference)."""
  it = test_dataset
  first = it()
  if callable(first):               # factory -> fresh iterator for this pass
    it, first = first, first()
  total, n, test_iter, inputs = 0.0, 0, 0, first
  while inputs is not None:
    t0 = time.time()
    _, loss, _ = model.build(inputs, False, costs_collection='costs')
    loss_value = float(loss.item())
    duration = time.time() - t0
    if world_info()[0] == 0:
      print('test: %s: step %d, loss = %.4f (%.1f examples/sec) %.3f sec/batch'
            % (datetime.now(), step, loss_value, opts['batch_size'] / float(duration), duration))
    total, n, test_iter = total + loss_value, n + 1, test_iter + 1
    try:
      inputs = it()
    except StopIteration:
      inputs = None
  print('iteration through test set finished')
  mean_loss = total / max(n, 1)
  if summary is not None and n:
    summary.writer.add_scalar('test/loss', mean_loss, step)
    summary.writer.flush()
  return mean_loss


def train_loop(opts, graph, loss, train_dataset, training_pl, handle_pl, train_op, train_summary_op,
               test_summary_op, num_steps, global_step, checkpoint_fname, test_dataset=None,
               ignore_missing_vars=False, reset_global_step=False, vars_to_restore=None, exclude_vars=None,
               fwd_only=False, allow_growth=False, model=None, log_every=1):
  """cnn_train_multi.py:371-516: restore, step loop with NaN guard and examples/sec logging, periodic checkpoints
  `<logdir>/model.ckpt-<step>.{index,data-00000-of-00001}` (TensorFlow TensorBundle format, the reference's variable names)."""
  rank = world_info()[0]
  eng = model.engine
  if checkpoint_fname and tf_checkpoint.checkpoint_exists(checkpoint_fname):          # cnn_train_multi.py:404
    print('RESTORING MODEL from: ' + checkpoint_fname)
    model.restore_checkpoint(checkpoint_fname, vars_to_restore=vars_to_restore or 'model',
                             ignore_missing_vars=ignore_missing_vars, exclude_vars=exclude_vars,
                             reset_global_step=reset_global_step if reset_global_step is not False else -1)
  start_step = int(eng.global_step) if eng is not None else -1
  summary = None
  if rank == 0 and opts.get('log_dir') and opts.get('n_summary') and not fwd_only and opts.get('summaries', True):
    from ..utils.summaries import SummaryLogger
    summary = SummaryLogger(opts['log_dir'])            # tf.summary.FileWriter(opts['log_dir']) (cnn_train_multi.py:436)
  begin = time.time()
  n_done = 0
  world = world_info()[2]
  for step in range(start_step, num_steps):
    t0 = time.time()
    if fwd_only:
      model.build(train_dataset(), False)
      loss_value = float(model.engine.loss_value().item())
    else:
      loss_value = float(train_op().item())          # the one D2H read per step (loss), as session.run returns it
      # the cost moving averages advance every step, as the reference's avg_ops inside train_op do (base_model.py:52-60;
      # rank-local values; the scalars were produced by this step's kernels and the loss read above already synchronised)
      for op in model._avg_ops:
        op()
    duration = time.time() - t0
    if world > 1 and dist.is_initialized() and step % log_every == 0:
      # the reference's loss is the tower MEAN (cnn_train_multi.py:177): average the rank-local values for the log line
      t = torch.tensor([loss_value], dtype=torch.float64, device=model.engine.dev)
      dist.all_reduce(t)
      loss_value = float(t.item()) / world
    assert not np.isnan(loss_value), 'Model diverged with loss = NaN'           # cnn_train_multi.py:463
    if rank == 0 and step % log_every == 0:
      print('%s: step %d, loss = %.4f (%.1f examples/sec) %.3f sec/batch'
            % (datetime.now(), step, loss_value, opts['batch_size'] / duration, duration))
    if summary is not None and step % opts['n_summary'] == 0:            # cnn_train_multi.py:452-457
      summary.write(model, step, lr=getattr(model.engine, 'last_lr', None), advance_avgs=fwd_only)
    if not fwd_only and test_dataset is not None and opts.get('n_test') and step % opts['n_test'] == 0:
      run_test_pass(opts, model, test_dataset, step, summary if rank == 0 else None)       # cnn_train_multi.py:472-506
    if not fwd_only and rank == 0 and step % opts['n_checkpoint'] == 0 and opts.get('log_dir'):
      # saver.save(session, <log_dir>/model.ckpt, global_step=step) (cnn_train_multi.py:511-513): TensorBundle files
      prefix = model.save_checkpoint(os.path.join(opts['log_dir'], 'model.ckpt-%d' % step))
      tf_checkpoint.update_checkpoint_state(opts['log_dir'], prefix)
    n_done += 1
  total = time.time() - begin
  if rank == 0 and n_done:
    print('Avg. samples per second %.3f' % (opts['batch_size'] * n_done / total))
