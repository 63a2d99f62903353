"""Quick device-side timing of the training step (development aid; bench.py is the contract)."""
import sys, time, torch
sys.path.insert(0, '.')
from imm_b200 import _lib
from imm_b200.engine import IMMEngine
from imm_b200.utils.box import default_model_config
from imm_b200.utils import synthetic as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 3
engine = int(sys.argv[3]) if len(sys.argv) > 3 else 0
eng = IMMEngine(default_model_config(10), B, 128, 'cuda:0', precision=prec, engine=engine)
eng.init_parameters(0); eng.load_vgg_caffe_dict(S.synthetic_vgg_caffe_dict(1))
inp = {k: v.cuda() for k, v in S.synthetic_inputs(B).items()}
print({k: v for k, v in eng.engine_table().items() if v != (2, 2, 2)})
for _ in range(3): eng.train_step(inp['image'], inp['future_image'], inp['mask'])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n0 = _lib.launch_count(); t0 = time.time(); e0.record()
import os
K = int(os.environ.get('IMMB_TS_STEPS', '20'))
for _ in range(K): eng.train_step(inp['image'], inp['future_image'], inp['mask'])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print('B=%d prec=%d: %.2f ms/step  %.1f pairs/s  %.1f TFLOP/s algorithmic; launches/step %d; wall %.2f ms; loss %.3f'
      % (B, prec, ms, B / ms * 1e3, B * 48.98e9 / ms / 1e9, (_lib.launch_count() - n0) // K if not eng.graph_replays else eng.graph_launches_per_step, (time.time() - t0) / K * 1e3, float(eng.total_loss)))
def timed(fn, n=3):
  torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n): fn()
  b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
print('fwd %.2f ms' % timed(lambda: eng.forward(inp['image'], inp['future_image'], inp['mask'])))
print('bwd %.2f ms' % timed(lambda: eng.backward()))
print('opt %.2f ms' % timed(lambda: eng.optimizer_step()))

# per-call attribution (CUDA events around every C-ABI call), single-stream schedule so that calls do not overlap
eng.wgrad_stream = eng.pose_stream = eng.gt_stream = None
eng.use_graph = False
_lib.PROFILE = []
eng.train_step(inp['image'], inp['future_image'], inp['mask'])
torch.cuda.synchronize()
import collections
by = collections.defaultdict(float); by_k = collections.defaultdict(float)
for name, tag, a, b, _info in _lib.PROFILE:
  ms_ = a.elapsed_time(b); by[(tag, name)] += ms_; by_k[name] += ms_
_lib.PROFILE = None
tot = sum(by.values())
print('sum of per-call times %.2f ms' % tot)
for (tag, name), v in sorted(by.items(), key=lambda kv: -kv[1])[:int(sys.argv[4]) if len(sys.argv) > 4 else 45]:
  print('  %-34s %-28s %7.3f ms' % (tag, name.replace('immb_', ''), v))
print('by kernel family:')
for name, v in sorted(by_k.items(), key=lambda kv: -kv[1])[:14]:
  print('  %-28s %7.3f ms' % (name.replace('immb_', ''), v))
