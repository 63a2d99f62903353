"""Development aid: BASELINE configs 3-5 shapes through the graph-replayed, multi-stream train step vs the eager
single-stream step (same seeds): prints ms/step and the loss trajectories.  usage: python tools/check_configs.py"""
import sys, torch
sys.path.insert(0, '.')
from imm_b200.engine import IMMEngine
from imm_b200.utils.box import default_model_config
from imm_b200.utils import synthetic as S

for name, K, B, R in (('config3 K=30 32/GPU', 30, 32, 128), ('config4 K=50 32/GPU', 50, 32, 128), ('config5 256px 16/GPU-sample', 10, 16, 256)):
  out = []
  for graph, streams in ((False, 0), (True, 3)):
    eng = IMMEngine(default_model_config(K), B, R, 'cuda:0', use_graph=graph, streams=streams)
    eng.init_parameters(0); eng.load_vgg_caffe_dict(S.synthetic_vgg_caffe_dict(1))
    inp = [{k: v.cuda() for k, v in S.synthetic_inputs(B, R, seed=i).items()} for i in range(2)]
    losses = []
    for i in range(4):
      d = inp[i % 2]
      losses.append(float(eng.train_step(d['image'], d['future_image'], d['mask'], lr=1e-3).item()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(6):
      d = inp[i % 2]
      eng.train_step(d['image'], d['future_image'], d['mask'], lr=1e-3)
    e1.record(); torch.cuda.synchronize()
    out.append((e0.elapsed_time(e1) / 6, losses, eng._graphs is not None))
    del eng
    torch.cuda.empty_cache()
  (t0, l0, g0), (t1, l1, g1) = out
  rel = max(abs(a - b) / abs(a) for a, b in zip(l0, l1))
  print('%-28s eager/1-stream %.2f ms | graph/3-stream %.2f ms (%s) -> %.0f pairs/s | loss rel diff %.1e  %s'
        % (name, t0, t1, g1, B / t1 * 1e3, rel, ['%.1f' % x for x in l1]))
