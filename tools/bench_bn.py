"""Micro-benchmark of the BN kernels (development aid)."""
import sys, torch
sys.path.insert(0, '.')
from imm_b200 import _lib
from imm_b200._lib import call
ST = _lib.stream_ptr
def timeit(fn, n=20):
  for _ in range(3): fn()
  torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n): fn()
  b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n * 1e3
for (N, H, C) in [(64, 128, 32), (64, 64, 64), (64, 32, 128), (64, 16, 256)]:
  npix = N * H * H
  y = torch.randn(npix, C, device='cuda'); g = torch.randn(npix, C, device='cuda')
  sums = torch.zeros(2 * C, dtype=torch.float64, device='cuda'); acc = torch.zeros(C, dtype=torch.float64, device='cuda')
  sc, sh, mu, inv = (torch.rand(C, device='cuda') + 0.5 for _ in range(4))
  dyh, dyl = torch.empty_like(y), torch.empty_like(y); dg, db = torch.empty(C, device='cuda'), torch.empty(C, device='cuda')
  oh, ol = torch.empty_like(y), torch.empty_like(y)
  scr = torch.empty(int(call('immb_bn_scratch_elems', npix, C)), dtype=torch.float64, device='cuda')
  flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
  t_stats = timeit(lambda: (flush.zero_(), call('immb_bn_stats', y, npix, C, C, sums, scr, scr.numel(), ST())))
  t_flush = timeit(lambda: flush.zero_())
  t_red = timeit(lambda: (flush.zero_(), call('immb_bn_bwd_reduce', g, C, y, C, npix, C, sc, sh, mu, inv, 1, sums, scr, scr.numel(), ST())))
  t_app = timeit(lambda: (flush.zero_(), call('immb_bn_bwd_apply', g, C, y, C, npix, C, sc, sh, mu, inv, 1, sums, dyh, dyl, dg, db, acc, scr, scr.numel(), ST())))
  t_fwd = timeit(lambda: (flush.zero_(), call('immb_bn_apply', y, N, H, H, C, C, sc, sh, 1, 0, oh, ol, C, ST())))
  gb = npix * C * 4 / 1e9
  print('N=%d H=%d C=%d (%.2f GB/tensor): stats %.0f us (%.2f TB/s)  bwd_reduce %.0f us (%.2f TB/s)  bwd_apply %.0f us (%.2f TB/s)  apply %.0f us (%.2f TB/s)'
        % (N, H, C, gb, t_stats - t_flush, gb / (t_stats - t_flush) * 1e3, t_red - t_flush, 2 * gb / (t_red - t_flush) * 1e3,
           t_app - t_flush, 4 * gb / (t_app - t_flush) * 1e3, t_fwd - t_flush, 3 * gb / (t_fwd - t_flush) * 1e3))
