import os, sys, torch
sys.path.insert(0, '/root/repo')
from imm_b200 import _lib
from imm_b200._lib import call
from tests.test_gpu_ops import conv_desc, split
def run(N,H,W,Cin,Cout,k,stride, prec):
    d = conv_desc(N,H,W,Cin,Cout,k,stride, engine=_lib.ENGINE_TC, precision=prec)
    x = torch.ones(N,H,W,Cin, device='cuda'); gy = torch.ones(N,d.Ho,d.Wo,Cout, device='cuda')
    z = torch.zeros_like(x); zg = torch.zeros_like(gy)
    dw = torch.full((k,k,Cin,Cout), float('nan'), device='cuda')
    ws = torch.empty(16, dtype=torch.uint8, device='cuda')
    call('immb_conv2d_wgrad', d, x, z, gy, zg, dw, ws, 16, _lib.stream_ptr()); torch.cuda.synchronize()
    print('dbg', os.environ.get('IMMB_WG_DBG'), (N,H,W,Cin,Cout,k,stride,prec), 'sum', float(dw.sum()), 'nnz', int((dw!=0).sum()), 'center', dw[k//2,k//2,0,:4].tolist(), 'corner', dw[0,0,0,:2].tolist(), flush=True)
run(1,16,16,32,32,3,1,1)
run(1,16,16,32,32,3,1,0)
run(2,32,32,128,128,3,1,1)
