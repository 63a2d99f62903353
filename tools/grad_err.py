"""Development aid: worst gradient-parity ratios (error / tolerance, tests/test_gpu_step_parity.py definition) of one
training step against the fp64 oracle.  usage: python tools/grad_err.py [batch] [n_maps]"""
import sys
sys.path.insert(0, '.')
import torch
from oracle import imm_oracle as O
from tests.gpu_util import make_pair, rel_err, to_dev
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
eng, st64, st32, inputs = make_pair(B, K, 128, 0)
r64 = O.train_step(st64, {k: v.double() for k, v in inputs.items()})
r32 = O.train_step(st32, inputs)
d = to_dev(inputs)
eng.train_step(d['image'], d['future_image'], d['mask'])
torch.cuda.synchronize()
worst = []
for k, g64 in r64['grads'].items():
  if k.endswith('/b') and ('/'.join(k.split('/')[:-2]) + '/batch_normalization/gamma') in r64['grads']:
    continue
  e_gpu, e_cpu = rel_err(eng.grads[k], g64), rel_err(r32['grads'][k], g64)
  worst.append((e_gpu / max(5.0 * e_cpu, 1e-2), k, e_gpu, e_cpu))
worst.sort(reverse=True)
for w in worst[:6]:
  print('ratio %.3f  %-64s gpu %.2e  cpu-fp32 %.2e' % w)
print('loss rel err %.2e' % (abs(float(eng.total_loss) - float(r64['loss'])) / float(r64['loss'])))
print('pred rel err %.2e' % rel_err(eng.pred[..., :3].cpu(), r64['out']['future_im_pred']))
