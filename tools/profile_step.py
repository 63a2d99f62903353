"""One training step bracketed by cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import sys, torch
sys.path.insert(0, '.')
from imm_b200.engine import IMMEngine
from imm_b200.utils.box import default_model_config
from imm_b200.utils import synthetic as S
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
eng = IMMEngine(default_model_config(K), B, 128, 'cuda:0')
eng.init_parameters(0); eng.load_vgg_caffe_dict(S.synthetic_vgg_caffe_dict(1))
inp = {k: v.cuda() for k, v in S.synthetic_inputs(B).items()}
for _ in range(2): eng.train_step(inp['image'], inp['future_image'], inp['mask'])
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.train_step(inp['image'], inp['future_image'], inp['mask'])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
