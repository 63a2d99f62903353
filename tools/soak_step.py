"""Soak the training step: N graph-replayed steps with a check after every step that no fp16-plane tensor saturated
(scale records) and that the loss is finite; reports the first offending step and the records involved.
usage: python tools/soak_step.py [steps] [batch]"""
import sys, torch
sys.path.insert(0, '.')
from imm_b200.engine import IMMEngine
from imm_b200.utils.box import default_model_config
from imm_b200.utils import synthetic as S
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng = IMMEngine(default_model_config(10), B, 128, 'cuda:0')
eng.init_parameters(0); eng.load_vgg_caffe_dict(S.synthetic_vgg_caffe_dict(1))
batches = [{k: v.cuda() for k, v in S.synthetic_inputs(B, seed=i).items()} for i in range(2)]
n = eng.n_scale_recs
bad_step = None
losses = []
for step in range(steps):
  d = batches[step % 2]
  loss = eng.train_step(d['image'], d['future_image'], d['mask'])
  torch.cuda.synchronize()
  recs = eng.scale_recs[:n].cpu()
  amax = recs[:, 1].contiguous().view(torch.float32)
  e = recs[:, 0].double()
  over = (~torch.isfinite(amax)) | (amax.double() * torch.exp2(e) > 65504.0)
  ov = int(eng.h16_overflow.item())
  lv = float(loss.item())
  losses.append(lv)
  if ov or bool(over.any()) or lv != lv:
    idx = [int(i) for i in over.nonzero().flatten()]
    print('step %d: overflow counter %d, loss %r, saturated records %s' % (step, ov, lv, [(i, eng.scale_tags[i], float(amax[i]), int(e[i])) for i in idx][:12]))
    bad_step = step
    break
print('soak: %d steps, first bad step %s, loss first %.3f last %.3f min %.3f max %.3f' %
      (len(losses), bad_step, losses[0], losses[-1], min(losses), max(l for l in losses if l == l)))
