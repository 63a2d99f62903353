#!/usr/bin/env python
"""Benchmark of the IMM training hot path (BASELINE.json metric: image-pairs/sec, 128x128, K=10).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...   # the restated reference on the host CPU cores

A "step" = one pass of the hot path over one batch of synthetic image pairs: forward + perceptual loss + backward
+ (N>1: one NCCL all-reduce of the flat gradient buffer) + per-tensor clip + Adam + BN/normaliser updates.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definition of every field."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BATCH = 64          # BASELINE.json configs[1]: CelebA-10pts, batch 64, 128x128, 1 x B200 (weak scaling: fixed per GPU)
N_MAPS = 10
IMAGE_SIZE = 128
GFLOP_PER_PAIR = 48.98      # algorithmic 2*MACs per image pair per training step (R=128, K=10)
GFLOP_VGG_PER_PAIR = 29.05  # of which frozen VGG16 tower: 2 x 9.683 forward + 9.683 dgrad (cost table)
# tensor-core passes per algorithmic MAC: 3 on the trainable stack (3xTF32), 2 on the frozen tower (weights exactly TF32)
MMA_PASSES = (3.0 * (GFLOP_PER_PAIR - GFLOP_VGG_PER_PAIR) + 2.0 * GFLOP_VGG_PER_PAIR) / GFLOP_PER_PAIR
METRIC = 'image-pairs/sec (128x128, K=10), training step'


def load_peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    d = json.load(open(p))
    return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
            'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler(object):
  """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md).  The sampler is started
  before the warm-up steps (nvidia-smi needs a few hundred ms to come up) and only the rows whose timestamp falls inside
  [mark_begin, mark_end] are reported."""
  Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
       'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
       'clocks_event_reasons.sw_power_cap')

  def __init__(self, gpu_index=0):
    self.rows, self.proc, self.gpu = [], None, gpu_index
    self.t0 = self.t1 = None

  def start(self):
    try:
      self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                    '--format=csv,noheader,nounits', '-lms', '20'],
                                   stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True)
      self.t.start()
    except Exception:
      self.proc = None

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append((time.time(), [x.strip() for x in line.split(',')]))

  def mark_begin(self):
    self.t0 = time.time()

  def mark_end(self):
    self.t1 = time.time()

  def stop(self):
    if self.proc is None:
      return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
    time.sleep(0.05)
    self.proc.terminate()
    try:
      self.proc.wait(timeout=2)
    except Exception:
      self.proc.kill()
    inside = [r for t, r in self.rows if self.t0 is not None and self.t0 <= t <= (self.t1 or t)]
    window = 'timed region'
    if len(inside) < 3:           # very short timed regions: fall back to every sample taken under load (warm-up + timed)
      inside, window = [r for _, r in self.rows], 'warm-up + timed region'
    sm, mx, reasons, power = [], [], set(), []
    for r in inside:
      try:
        sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
        for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
          if val.lower().startswith('active'):
            reasons.add(name)
      except Exception:
        pass
    sm.sort()
    return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
            'power_w_max': max(power) if power else None, 'samples': len(sm), 'window': window,
            'reasons': sorted(reasons)}


# --------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_batch):
  """Times the CPU restatement of the reference (oracle/imm_oracle.py, PyTorch-CPU fp32, all host threads) on a
  bounded sample of the workload: `sample_batch` pairs per step instead of PER_GPU_BATCH.  TensorFlow 1.10 is not
  installable, so this is kind 'port' (labelled 'restated reference, not TF1')."""
  import torch
  from oracle import imm_oracle as O
  cores = os.cpu_count() or 1
  st = O.init_state(O.State(n_maps=N_MAPS, image_size=IMAGE_SIZE), seed=0)
  inp = O.synthetic_inputs(sample_batch, IMAGE_SIZE, seed=0)
  # "all the host threads it can use": oneDNN/OpenMP scale badly past a few dozen threads at this problem size,
  # so time one step per candidate thread count and keep the fastest (reported as `cores`)
  best, best_t = cores, None
  for n in sorted({min(cores, c) for c in (16, 32, 64)}):
    torch.set_num_threads(n)
    O.train_step(st, inp)
    t0 = time.time()
    O.train_step(st, inp)
    dt = time.time() - t0
    if best_t is None or dt < best_t:
      best, best_t = n, dt
  cores = best
  torch.set_num_threads(cores)
  for _ in range(warmup):
    O.train_step(st, inp)
  t0 = time.time()
  for _ in range(steps):
    O.train_step(st, inp)
  dt = (time.time() - t0) / max(steps, 1)
  return {'pairs_per_s': sample_batch / dt, 'ms_per_step': dt * 1e3, 'cores': cores,
          'sample': '%d steps of %d pairs (batch %d is the workload; CPU step time scales linearly in pairs)'
                    % (steps, sample_batch, PER_GPU_BATCH)}


def workload_config(n_gpus):
  return {'workload': 'CelebA-10pts model section, batch %d per GPU (global %d), 128x128x3 synthetic image pairs + '
                      'reference smooth mask, full train step (fwd, VGG16 perceptual loss, bwd, clip, Adam)'
                      % (PER_GPU_BATCH, PER_GPU_BATCH * n_gpus),
          'global_batch': PER_GPU_BATCH * n_gpus, 'image_size': IMAGE_SIZE, 'n_maps': N_MAPS,
          'parallelism': 'dp%d' % n_gpus,
          'weights': 'reference initialisers (seeded); VGG16: seeded synthetic weights in the Caffe-dict layout',
          'l2': 'per-step working set (several GB of activations at batch 64) exceeds the 126 MB L2; no explicit flush'}


def main_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  sample_batch = 8
  r = cpu_reference_run(args.steps, args.warmup, sample_batch)
  line = {'impl': 'reference', 'metric': METRIC, 'value': r['pairs_per_s'], 'unit': 'pairs/s', 'n_gpus': args.gpus,
          'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True,
          'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': workload_config(args.gpus),
          'cpu_baseline': {'value': r['pairs_per_s'], 'unit': 'pairs/s', 'cores': r['cores'], 'kind': 'port',
                           'sample': r['sample'] + '; restated reference (PyTorch-CPU fp32), not TF1'},
          'e2e': {'value': r['pairs_per_s'], 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
  print(json.dumps(line))


def main_cuda(args):
  import torch
  import torch.distributed as dist
  from imm_b200 import _lib
  from imm_b200.models.imm_model import IMMModel
  from imm_b200.train import cnn_train_multi as tru
  from imm_b200.utils.box import default_model_config
  from imm_b200.utils.synthetic import synthetic_inputs, synthetic_vgg_caffe_dict

  if not torch.cuda.is_available():
    raise SystemExit('bench.py: no CUDA device (the CUDA path has no CPU fallback)')
  rank, local_rank, world = tru.init_distributed('nccl')
  torch.cuda.set_device(local_rank)
  dev = 'cuda:%d' % local_rank
  B = PER_GPU_BATCH
  model = IMMModel(default_model_config(N_MAPS), global_step=-1, device=dev, world_size=world,
                   vgg_data=synthetic_vgg_caffe_dict(1), seed=0)
  optim = tru.AdamOptimizer(tru.exponential_decay(1e-3, 100000, 0.95))
  host = [synthetic_inputs(B, IMAGE_SIZE, seed=rank * 10 + i, pin=True) for i in range(2)]
  resident = [{k: v.to(dev) for k, v in h.items()} for h in host]
  allreduce = tru.average_gradients if world > 1 else None

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def max_over_ranks(ms):
    if world == 1:
      return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

  # ---- kernel-resident arm: inputs already in HBM ---------------------------------------------------------
  model.build(resident[0], True)            # instantiates the engine
  eng = model.engine

  def step_resident(i, eager=False):
    d = resident[i % 2]
    if eager:       # per-call profiling step: explicit phases, no graph replay
      eng.forward(d['image'], d['future_image'], d['mask'], training=True, build_loss=True)
      eng.backward()
      eng.optimizer_step(1.0, lr=optim.lr(eng.global_step), allreduce=allreduce)
    else:           # the engine's train step (captured into CUDA graphs after two eager warm-up steps)
      eng.train_step(d['image'], d['future_image'], d['mask'], clip_value=1.0, lr=optim.lr(eng.global_step),
                     allreduce=allreduce)

  sampler = ClockSampler(local_rank)
  if rank == 0:
    sampler.start()
  for i in range(args.warmup):
    step_resident(i)
  barrier()
  sampler.mark_begin()
  n0, r0 = _lib.launch_count(), eng.graph_replays
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for i in range(args.steps):
    step_resident(i)
  e1.record()
  barrier()
  sampler.mark_end()
  # kernels launched in the timed region: eager launches + (graph replays x kernels recorded per graph)
  launches = _lib.launch_count() - n0 + (eng.graph_replays - r0) * eng.graph_launches_per_step
  ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
  clocks = sampler.stop() if rank == 0 else None
  value = B * world / (ms * 1e-3)

  # ---- end-to-end arm: public API, pinned host inputs, H2D inside the timed region, loss read back -----------
  train_op_inputs = {'i': 0}

  def next_host():
    b = host[train_op_inputs['i'] % 2]
    train_op_inputs['i'] += 1
    return b
  _, train_op, _, _, _ = tru.setup_training({'gpu_ids': list(range(world)), 'batch_size': B * world}, None, optim,
                                            next_host, True, type('F', (), {'create': staticmethod(lambda: model)}),
                                            -1, clip_value=1.0)
  for _ in range(max(args.warmup, 3)):
    float(train_op().item())
  barrier()
  e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e2.record()
  for _ in range(args.steps):
    loss_val = float(train_op().item())
  e3.record()
  barrier()
  ms_e2e = max_over_ranks(e2.elapsed_time(e3)) / args.steps
  h2d = sum(v.numel() * v.element_size() for v in host[0].values()) * world
  e2e = {'value': B * world / (ms_e2e * 1e-3), 'unit': 'pairs/s', 'ms_per_step': ms_e2e,
         'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4 * world,
         'api': 'setup_training() train_op on IMMModel (inputs in pinned host memory, one-deep H2D prefetch on a copy stream, '
                'CUDA-graph replay of the step) + loss.item() per step'}

  # ---- roofline of the dominant kernel family: one extra step with CUDA events around every C-ABI call --------
  # (single-stream schedule for this one step, so that per-call durations are not inflated by overlapping kernels)
  saved_streams = (eng.wgrad_stream, eng.pose_stream, eng.gt_stream)
  eng.wgrad_stream = eng.pose_stream = eng.gt_stream = None
  _lib.PROFILE = []
  step_resident(0, eager=True)
  torch.cuda.synchronize()
  eng.wgrad_stream, eng.pose_stream, eng.gt_stream = saved_streams
  fam, kern = {}, {}
  for name, tag, a, b, info in _lib.PROFILE:
    t = a.elapsed_time(b)
    fam[name] = fam.get(name, 0.0) + t
    if info is not None:
      k = kern.setdefault(info['kernel'], {'ms': 0.0, 'flops': 0.0, 'mma_flops': 0.0, 'launches': 0})
      k['ms'] += t
      k['flops'] += info['flops']
      k['mma_flops'] += info['flops'] * info['passes']
      k['launches'] += 1
  _lib.PROFILE = None
  conv_ms = sum(k['ms'] for k in kern.values())
  peaks = load_peaks()
  conv_tflops = GFLOP_PER_PAIR * B / conv_ms            # GFLOP / ms == TFLOP/s
  top_name = max(kern, key=lambda n: kern[n]['ms'])
  top = kern[top_name]
  top_tflops = top['flops'] / top['ms'] * 1e-9          # algorithmic FLOP per launch / average launch duration
  top_mma_tflops = top['mma_flops'] / top['ms'] * 1e-9
  traffic = None
  tj = os.path.join(ROOT, 'profiles', 'top_kernel.json')
  if os.path.exists(tj):
    traffic = json.load(open(tj)).get('dram_bytes_per_launch')
  tf32_peak = 0.5 * peaks['bf16_tflops_sustained']
  roofline = {'bound': 'tensor',
              'kernel': '%s (persistent CTA-pair halo conv: tcgen05 cta_group::2 kind::tf32, M=256, TMA halo boxes; forward and '
                        'dgrad of every stride-1 3x3 layer and the 7x7 first layer) -- %d launches, %.1f%% of the step'
                        % (top_name, top['launches'], 100.0 * top['ms'] / ms),
              'achieved': top_tflops, 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
              'frac': top_tflops / peaks['bf16_tflops_sustained'], 'traffic': traffic,
              'algorithmic_gflop_per_launch': top['flops'] / top['launches'] * 1e-9,
              'ms_per_launch': top['ms'] / top['launches'],
              'peak_source': 'MEASURED_PEAKS.json bf16 sustained (%s).  The kernel runs kind::tf32 (half the bf16 rate) and issues '
                             '3 MMAs per algorithmic MAC on the trainable stack (error-compensated 3xTF32) or 2 on the frozen VGG '
                             'tower, so `frac` is capped near %.3f; `tensor_pipe_frac_tf32` = issued TF32 MMA FLOP/s / (peak/2) is '
                             'the issue-rate view of the same thing; the sustained bf16 GEMM that sets `peak` runs power-capped near 1.3 GHz while this '
                             'kernel holds 1.85-1.97 GHz, so the ratio can exceed ncu sm__pipe_tensor_cycles_active (70.6 %% '
                             'time-weighted over 21 launches, profiles/top_kernel.json), which is the utilisation figure'
                             % (peaks['source'], 0.5 / MMA_PASSES),
              'tensor_pipe_frac_tf32': top_mma_tflops / tf32_peak,
              'conv_engine': {'flops_accounted_vs_survey': sum(k['flops'] for k in kern.values()) * 1e-9 / (GFLOP_PER_PAIR * B),
                              'ms_per_step': conv_ms, 'step_share': conv_ms / ms, 'achieved_tflops_algorithmic': conv_tflops,
                              'tensor_pipe_frac_tf32': MMA_PASSES * conv_tflops / tf32_peak,
                              'by_kernel': {n: {'ms': round(k['ms'], 3), 'launches': k['launches'],
                                                'tflops_algorithmic': round(k['flops'] / k['ms'] * 1e-9, 1),
                                                'tensor_pipe_frac_tf32': round(k['mma_flops'] / k['ms'] * 1e-9 / tf32_peak, 3)}
                                            for n, k in sorted(kern.items(), key=lambda kv: -kv[1]['ms'])}},
              'per_family_ms': {k.replace('immb_', ''): round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])[:10]}}

  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  if rank != 0:
    return
  cpu = None
  if world == 1 and not args.no_cpu_baseline:
    r = cpu_reference_run(3, 1, 8)
    cpu = {'value': r['pairs_per_s'], 'unit': 'pairs/s', 'cores': r['cores'], 'kind': 'port',
           'sample': r['sample'] + '; restated reference (PyTorch-CPU fp32), not TF1'}
  line = {'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
          'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
          'dtype': 'tf32x3 (fp32 storage; error-compensated TF32 tensor-core products hi*hi+hi*lo+lo*hi, fp32 accumulate; '
                   'the frozen VGG16 tower uses weights rounded to TF32 at load and 2 passes)',
          'data': 'synthetic', 'config': workload_config(world), 'tflops_algorithmic': GFLOP_PER_PAIR * B * world / ms,
          'roofline': roofline, 'cpu_baseline': cpu, 'clocks': clocks, 'e2e': e2e,
          'gpu_launches': int(launches) * world, 'last_loss': loss_val,
          'streams': {'wgrad_side_stream': eng.wgrad_stream is not None, 'pose_branch_stream': eng.pose_stream is not None, 'vgg_gt_half_stream': eng.gt_stream is not None,
                      'input_prefetch_stream': True, 'cuda_graph_replay': eng._graphs is not None}}
  print(json.dumps(line))
  sys.stdout.flush()


if __name__ == '__main__':
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', type=str, default='cuda', choices=['cuda', 'reference'])
  ap.add_argument('--no-cpu-baseline', action='store_true')
  a = ap.parse_args()
  a.warmup = max(a.warmup, 3) if a.impl == 'cuda' else a.warmup
  if a.impl == 'reference':
    main_reference(a)
  else:
    main_cuda(a)
