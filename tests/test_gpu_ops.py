"""Kernel-level parity tests: each C-ABI op against the oracle's restatement of the TF op it replaces
(fp64 on CPU, autograd for the backward ops).  Edge cases: odd channel counts, channel-strided concat
buffers, stride-2 asymmetric SAME padding, 1x1 / 7x7 kernels, tiny and ragged sizes."""
import math

import numpy as np
import pytest
import torch

from imm_b200 import _lib
from imm_b200._lib import ConvDesc, call
from oracle import imm_oracle as O
from tests.gpu_util import rel_err

pytestmark = pytest.mark.gpu
ST = lambda: _lib.stream_ptr()


def split(t):
  """host-side TF32 split (round-to-nearest-even vs the kernel's rna only differ on exact ties)."""
  hi = O.round_tf32(t)
  lo = O.round_tf32(t - hi)
  return hi, lo


def conv_desc(N, H, W, Cin, Cout, k, stride, xcs=None, engine=_lib.ENGINE_AUTO, epilogue=0, precision=0, ycs=None):
  d = ConvDesc()
  d.N, d.H, d.W, d.Cin, d.Cout, d.kh, d.kw, d.stride = N, H, W, Cin, Cout, k, k, stride
  d.Ho, d.Wo = -(-H // stride), -(-W // stride)
  d.pad_t, d.pad_l = O.same_pad(H, k, stride)[0], O.same_pad(W, k, stride)[0]
  d.x_cstride = xcs or Cin
  d.y_cstride = ycs or Cout
  d.cin_pad = (Cin + 31) // 32 * 32
  d.epilogue, d.precision, d.engine = epilogue, precision, engine
  return d


CONV_CASES = [
  # N, H, W, Cin, Cout, k, stride, xcs
  (2, 16, 16, 3, 32, 7, 1, None),        # encoder conv_1 (7x7, Cin=3)
  (1, 12, 20, 1, 64, 3, 1, None),        # vgg conv1_1 (Cin=1), non-square
  (2, 16, 16, 32, 64, 3, 2, None),       # stride-2, asymmetric SAME pad
  (1, 9, 7, 5, 6, 3, 2, None),           # odd sizes, stride 2 (pad 1/1), ragged tiles
  (2, 16, 16, 266, 256, 3, 1, 288),      # renderer conv_1: concat input with padded channel stride
  (2, 16, 16, 256, 10, 1, 1, None),      # pose 1x1 -> K heatmaps
  (1, 32, 32, 32, 9, 3, 1, None),        # renderer last conv (Cout=9)
  (3, 8, 8, 64, 64, 3, 1, None),
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_fwd_dgrad_wgrad_simt(case):
  N, H, W, Cin, Cout, k, stride, xcs = case
  g = torch.Generator().manual_seed(hash(case) % 1000)
  xcs_ = xcs or Cin
  x = torch.randn(N, H, W, xcs_, generator=g)
  x[..., Cin:] = 0
  w = torch.randn(k, k, Cin, Cout, generator=g) * 0.1
  b = torch.randn(Cout, generator=g)
  d = conv_desc(N, H, W, Cin, Cout, k, stride, xcs, engine=_lib.ENGINE_SIMT)
  xd = x.double()[..., :Cin].clone().requires_grad_(True)
  wd = w.double().clone().requires_grad_(True)
  y_ref = O.conv2d_same(xd, wd, b.double(), stride)
  gy = torch.randn(y_ref.shape, generator=g)
  y_ref.backward(gy.double())
  xh, xl = split(x)
  y = torch.empty(N, d.Ho, d.Wo, Cout, device='cuda')
  call('immb_conv2d_fwd', d, xh.cuda(), xl.cuda(), w.cuda(), None, None, b.cuda(), y, None, ST())
  assert rel_err(y, y_ref) < 1e-5
  gh, gl = split(gy)
  dx = torch.zeros(N, H, W, xcs_, device='cuda')
  call('immb_conv2d_dgrad', d, gh.cuda(), gl.cuda(), w.cuda(), None, None, dx, ST())
  assert rel_err(dx[..., :Cin], xd.grad) < 1e-5
  dw = torch.empty(k, k, Cin, Cout, device='cuda')
  ws = torch.empty(16, dtype=torch.uint8, device='cuda')
  call('immb_conv2d_wgrad', d, xh.cuda(), xl.cuda(), gh.cuda(), gl.cuda(), dw, ws, 16, ST())
  assert rel_err(dw, wd.grad) < 1e-5
  # fused bias+ReLU epilogue writing split planes (VGG path)
  d2 = conv_desc(N, H, W, Cin, Cout, k, stride, xcs, engine=_lib.ENGINE_SIMT, epilogue=_lib.EPI_BIAS_RELU)
  yh, yl = torch.empty_like(y), torch.empty_like(y)
  call('immb_conv2d_fwd', d2, xh.cuda(), xl.cuda(), w.cuda(), None, None, b.cuda(), yh, yl, ST())
  assert rel_err(yh + yl, torch.relu(y_ref)) < 1e-5
  assert float((yh.cpu() - O.round_tf32(yh.cpu())).abs().max()) == 0.0      # hi plane is exactly TF32


TC_CASES = [
  # N, H, W, Cin, Cout, k, stride, xcs
  (2, 16, 16, 32, 32, 3, 1, None),
  (1, 32, 32, 64, 128, 3, 1, None),
  (2, 16, 16, 64, 64, 1, 1, None),
  (2, 32, 32, 32, 64, 3, 2, None),       # stride 2: parity-split TMA view, 4-class dgrad
  (2, 64, 64, 64, 128, 3, 2, None),
  (2, 16, 16, 266, 256, 3, 1, 288),      # renderer conv_1: 266 logical channels, stride 288, BN=96 dgrad
  (4, 8, 8, 128, 128, 3, 1, None),       # 8x8 maps: 2 images per 128-pixel tile
  (3, 8, 8, 128, 64, 3, 1, None),        # odd image count -> predicated tile overhang
  (1, 128, 128, 32, 32, 3, 1, None),     # encoder conv_2 shape
  (1, 32, 32, 64, 32, 3, 1, None),
  (2, 16, 16, 512, 512, 3, 1, None),     # VGG conv4_x
  (1, 32, 32, 32, 9, 3, 1, None),        # renderer last conv: Cout=9 in a 12-channel-stride tensor (BN=16 tile)
  (2, 16, 16, 256, 10, 1, 1, None),      # pose 1x1 -> 10 heatmaps (stride 12)
  (1, 16, 16, 256, 50, 1, 1, None),      # pose 1x1 -> 50 heatmaps (stride 52, BN=64 tile)
  (2, 32, 32, 9, 64, 1, 1, 12),          # vgg conv1_1 as a 1x1 conv over 3x3 patches (Cin=9, stride 12)
  (3, 16, 16, 64, 96, 3, 1, None),       # CTA-pair N tile of 96 (48 weight rows per CTA), odd image count
  (1, 16, 16, 320, 192, 3, 1, None),     # pair N tiles of 192 (fwd) and 2 x 160 (dgrad over 320 input channels)
]


@pytest.mark.parametrize('precision', [_lib.PREC_TF32X3, _lib.PREC_TF32])
@pytest.mark.parametrize('case', TC_CASES)
def test_conv_tcgen05_engine(case, precision):
  """tcgen05/TMA engine vs the fp64 oracle.  3xTF32 must be at fp32 accuracy; single-pass TF32 at ~1e-3."""
  N, H, W, Cin, Cout, k, stride, xcs = case
  g = torch.Generator().manual_seed(sum(case[:7]))
  xcs_ = xcs or Cin
  x = torch.randn(N, H, W, xcs_, generator=g)
  x[..., Cin:] = 0
  w = torch.randn(k, k, Cin, Cout, generator=g) * 0.1
  b = torch.randn(Cout, generator=g)
  ycs = (Cout + 3) // 4 * 4
  d = conv_desc(N, H, W, Cin, Cout, k, stride, xcs, engine=_lib.ENGINE_TC, precision=precision, ycs=ycs)
  assert [_lib.lib().immb_conv_engine_for(d, op) for op in range(3)] == [_lib.ENGINE_TC] * 3
  # tensor-core fp32 accumulation truncates (round-toward-zero) at every K=8 step: the error grows with the
  # reduction length (K = 4608 for the 512-channel VGG layers) but stays ~30x below the 1e-3 bar
  tol = (2e-5 if k * k * Cin < 2048 else 6e-5) if precision == _lib.PREC_TF32X3 else 3e-3
  xd = x.double()[..., :Cin].clone().requires_grad_(True)
  wd = w.double().clone().requires_grad_(True)
  y_ref = O.conv2d_same(xd, wd, b.double(), stride)
  gy = torch.randn(y_ref.shape, generator=g)
  y_ref.backward(gy.double())
  dev = 'cuda'
  taps, cp = k * k, d.cin_pad
  wp_h, wp_l = torch.empty(taps, Cout, cp, device=dev), torch.empty(taps, Cout, cp, device=dev)
  wh_h, wh_l = torch.empty(taps, cp, ycs, device=dev), torch.empty(taps, cp, ycs, device=dev)
  call('immb_pack_weights', w.to(dev), k, k, Cin, Cout, cp, ycs, wp_h, wp_l, wh_h, wh_l, ST())
  xh, xl = split(x)
  xh, xl = xh.to(dev), xl.to(dev)
  y = torch.full((N, d.Ho, d.Wo, ycs), float('nan'), device=dev)
  call('immb_conv2d_fwd', d, xh, xl, None, wp_h, wp_l, b.to(dev), y, None, ST())
  torch.cuda.synchronize()
  assert rel_err(y[..., :Cout], y_ref) < tol, ('fwd', rel_err(y[..., :Cout], y_ref))
  if ycs > Cout:
    assert float(y[..., Cout:].abs().max()) == 0.0
  gyp = torch.zeros(N, d.Ho, d.Wo, ycs)
  gyp[..., :Cout] = gy
  gh, gl = split(gyp)
  gh, gl = gh.to(dev), gl.to(dev)
  dx = torch.full((N, H, W, xcs_), float('nan'), device=dev)
  call('immb_conv2d_dgrad', d, gh, gl, None, wh_h, wh_l, dx, ST())
  torch.cuda.synchronize()
  assert rel_err(dx[..., :Cin], xd.grad) < tol, ('dgrad', rel_err(dx[..., :Cin], xd.grad))
  if xcs_ > Cin:
    assert float(dx[..., Cin:].abs().max()) == 0.0       # padded channels come out as exact zeros
  dw = torch.full((k, k, Cin, Cout), float('nan'), device=dev)
  ws = torch.empty(16, dtype=torch.uint8, device=dev)
  call('immb_conv2d_wgrad', d, xh, xl, gh, gl, dw, ws, 16, ST())
  torch.cuda.synchronize()
  assert rel_err(dw, wd.grad) < tol, ('wgrad', rel_err(dw, wd.grad))
  # fused bias+ReLU epilogue writing split planes (VGG path)
  d2 = conv_desc(N, H, W, Cin, Cout, k, stride, xcs, engine=_lib.ENGINE_TC, epilogue=_lib.EPI_BIAS_RELU,
                 precision=precision, ycs=ycs)
  yh, yl = torch.empty_like(y), torch.empty_like(y)
  call('immb_conv2d_fwd', d2, xh, xl, None, wp_h, wp_l, b.to(dev), yh, yl, ST())
  assert rel_err((yh + yl)[..., :Cout], torch.relu(y_ref)) < tol


def test_conv_tc2_cluster_multicast_mode():
  """IMMB_TC2_CLUSTER=2 forces the 2-CTA-cluster variant of the halo kernel (each CTA TMA-multicasts half of every
  weight slice into both CTAs' shared memory).  The mode is read once per process, so run a subset in a child."""
  import os, subprocess, sys
  env = dict(os.environ, IMMB_TC2_CLUSTER='2', IMMB_TC2_PAIR='0')
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  out = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_ops.py'), '-m', 'gpu', '-q',
                        '-k', 'tcgen05_engine and (case0 or case1 or case5 or case10)', '-p', 'no:cacheprovider'],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=root, timeout=600)
  assert out.returncode == 0 and 'passed' in out.stdout and 'failed' not in out.stdout, out.stdout[-1500:]


def test_conv_dgrad_fused_relu_backward():
  """immb_conv2d_dgrad_relu = dgrad * [act > 0] written as TF32 split planes (VGG backward: the producer's ReLU
  backward and the operand split folded into the consumer's dgrad epilogue) vs the two separate ops."""
  N, H, W, Cin, Cout = 2, 32, 32, 64, 128
  g = torch.Generator().manual_seed(21)
  w = torch.randn(3, 3, Cin, Cout, generator=g) * 0.1
  gy = torch.randn(N, H, W, Cout, generator=g)
  act = torch.relu(torch.randn(N, H, W, Cin, generator=g))          # post-ReLU activation that fed the conv
  dev = 'cuda'
  d = conv_desc(N, H, W, Cin, Cout, 3, 1, None, engine=_lib.ENGINE_TC, precision=_lib.PREC_TF32X2)
  assert _lib.lib().immb_conv2d_dgrad_relu_supported(d) == 1
  wq = O.round_tf32(w)                                               # 2-pass product: weights exactly TF32
  wp_h, wp_l = torch.empty(9, Cout, Cin, device=dev), torch.empty(9, Cout, Cin, device=dev)
  wh_h, wh_l = torch.empty(9, Cin, Cout, device=dev), torch.empty(9, Cin, Cout, device=dev)
  call('immb_pack_weights', wq.to(dev), 3, 3, Cin, Cout, Cin, Cout, wp_h, wp_l, wh_h, wh_l, ST())
  gh, gl = split(gy)
  gh, gl = gh.to(dev), gl.to(dev)
  dx = torch.empty(N, H, W, Cin, device=dev)
  call('immb_conv2d_dgrad', d, gh, gl, None, wh_h, wh_l, dx, ST())
  ah, _ = split(act)
  oh, ol = torch.full((N, H, W, Cin), float('nan'), device=dev), torch.full((N, H, W, Cin), float('nan'), device=dev)
  call('immb_conv2d_dgrad_relu', d, gh, gl, wh_h, wh_l, ah.to(dev), Cin, oh, ol, ST())
  torch.cuda.synchronize()
  want = dx.cpu() * (act > 0).float()
  got = (oh + ol).cpu()
  assert rel_err(got, want) < 1e-6                                   # same kernel, same accumulation; only the split differs
  assert float((oh.cpu() - O.round_tf32(oh.cpu())).abs().max()) == 0.0     # hi plane is exactly TF32
  assert float(got[act <= 0].abs().max()) == 0.0                      # the ReLU backward mask is exact
  xd = torch.zeros(N, H, W, Cin, dtype=torch.float64, requires_grad=True)
  O.conv2d_same(xd, wq.double(), None, 1).backward(gy.double())
  assert rel_err(oh + ol, xd.grad * (act > 0).double()) < 2e-5
  # shapes outside the halo pair kernel are refused, not silently routed elsewhere
  d8 = conv_desc(N, 8, 8, Cin, Cout, 3, 1, None, engine=_lib.ENGINE_TC, precision=_lib.PREC_TF32X2)
  assert _lib.lib().immb_conv2d_dgrad_relu_supported(d8) == 0


def test_conv_tc2_single_cta_mode():
  """IMMB_TC2_PAIR=0 routes the stride-1 3x3 layers through the single-CTA halo kernel (cta_group::1) instead of the
  default CTA-pair kernel (cta_group::2, M = 256); both must meet the same bars.  Read once per process -> child."""
  import os, subprocess, sys
  env = dict(os.environ, IMMB_TC2_PAIR='0')
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  out = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(root, 'tests', 'test_gpu_ops.py'), '-m', 'gpu', '-q',
                        '-k', 'tcgen05_engine and (case0 or case1 or case5 or case8 or case10 or case11 or case15)',
                        '-p', 'no:cacheprovider'],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=root, timeout=600)
  assert out.returncode == 0 and 'passed' in out.stdout and 'failed' not in out.stdout, out.stdout[-1500:]


def test_pack_weights_layouts():
  w = torch.randn(3, 3, 5, 7)
  wp_h, wp_l = torch.empty(9, 7, 32, device='cuda'), torch.empty(9, 7, 32, device='cuda')
  wh_h, wh_l = torch.empty(9, 32, 8, device='cuda'), torch.empty(9, 32, 8, device='cuda')
  call('immb_pack_weights', w.cuda(), 3, 3, 5, 7, 32, 8, wp_h, wp_l, wh_h, wh_l, ST())
  full = torch.zeros(9, 32, 8)
  full[:, :5, :7] = w.view(9, 5, 7)
  assert rel_err(wh_h + wh_l, full) < 1e-6
  assert rel_err(wp_h + wp_l, full[:, :, :7].permute(0, 2, 1)) < 1e-6
  assert float(wh_h[:, 5:].abs().max()) == 0.0 and float(wh_h[:, :, 7:].abs().max()) == 0.0


@pytest.mark.parametrize('shape', [(2, 8, 8, 32), (3, 5, 7, 48), (1, 16, 16, 256)])
@pytest.mark.parametrize('up2x', [0, 1])
def test_bn_train_fwd_bwd(shape, up2x):
  N, H, W, C = shape
  g = torch.Generator().manual_seed(N * 100 + C + up2x)
  y = torch.randn(shape, generator=g) * 3 + 5
  gamma = torch.rand(C, generator=g) + 0.5
  beta = torch.randn(C, generator=g) * 0.1
  mm0, mv0 = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
  yd = y.double().requires_grad_(True)
  gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
  z, mm_ref, mv_ref = O.batch_norm(yd, gd, bd, mm0.double(), mv0.double(), True)
  a = torch.relu(z)
  if up2x:
    a = O.resize_bilinear(a, [2 * H, 2 * W])
  go = torch.randn(a.shape, generator=g)
  a.backward(go.double())
  dev = 'cuda'
  yc = y.to(dev)
  sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
  mm, mv = mm0.to(dev), mv0.to(dev)
  scale, shift, mean, invstd = (torch.empty(C, device=dev) for _ in range(4))
  npix = N * H * W
  scr = torch.empty(int(call('immb_bn_scratch_elems', npix, C)) if up2x else 0, dtype=torch.float64, device=dev)   # both paths
  scr_p = scr if scr.numel() else None
  call('immb_bn_stats', yc, npix, C, C, sums, scr_p, scr.numel(), ST())
  call('immb_bn_finalize', sums, npix, C, gamma.to(dev), beta.to(dev), mm, mv, 1, scale, shift, mean, invstd, ST())
  s = 2 if up2x else 1
  ocs = C + 8       # write into a wider (concat-style) buffer
  oh = torch.zeros(N, H * s, W * s, ocs, device=dev)
  ol = torch.zeros_like(oh)
  call('immb_bn_apply', yc, N, H, W, C, C, scale, shift, 1, up2x, oh, ol, ocs, ST())
  assert rel_err((oh + ol)[..., :C], a) < 1e-5
  assert float((oh + ol)[..., C:].abs().max()) == 0.0
  assert rel_err(mm, mm_ref) < 1e-5 and rel_err(mv, mv_ref) < 1e-5
  # backward
  gdev = go.to(dev)
  if up2x:
    glow = torch.empty(N, H, W, C, device=dev)
    call('immb_upsample2x_bwd', gdev, N, H, W, C, C, glow, ST())
    gdev = glow
  bs = torch.zeros(2 * C, dtype=torch.float64, device=dev)
  dbacc = torch.zeros(C, dtype=torch.float64, device=dev)
  dyh, dyl = torch.empty(N, H, W, C, device=dev), torch.empty(N, H, W, C, device=dev)
  dg, db_ = torch.empty(C, device=dev), torch.empty(C, device=dev)
  call('immb_bn_bwd_reduce', gdev, C, yc, C, npix, C, scale, shift, mean, invstd, 1, bs, scr_p, scr.numel(), ST())
  call('immb_bn_bwd_apply', gdev, C, yc, C, npix, C, scale, shift, mean, invstd, 1, bs, dyh, dyl, dg, db_, dbacc, scr_p, scr.numel(), ST())
  assert rel_err(dyh + dyl, yd.grad) < 1e-4
  assert rel_err(dg, gd.grad) < 1e-4 and rel_err(db_, bd.grad) < 1e-4
  assert float(dbacc.abs().max()) < 1e-3 * float(yd.grad.abs().sum())     # sum(dy) is analytically 0


def test_bn_eval_uses_moving_stats():
  C = 16
  y = torch.randn(2, 4, 4, C)
  gamma, beta = torch.rand(C) + 0.5, torch.randn(C)
  mm, mv = torch.randn(C), torch.rand(C) + 0.5
  ref, _, _ = O.batch_norm(y.double(), gamma.double(), beta.double(), mm.double(), mv.double(), False)
  dev = 'cuda'
  mmd, mvd = mm.to(dev), mv.to(dev)
  scale, shift, mean, invstd = (torch.empty(C, device=dev) for _ in range(4))
  call('immb_bn_finalize', None, 32, C, gamma.to(dev), beta.to(dev), mmd, mvd, 0, scale, shift, mean, invstd, ST())
  oh, ol = torch.empty(2, 4, 4, C, device=dev), torch.empty(2, 4, 4, C, device=dev)
  call('immb_bn_apply', y.to(dev), 2, 4, 4, C, C, scale, shift, 0, 0, oh, ol, C, ST())
  assert rel_err(oh + ol, ref) < 1e-5
  assert torch.equal(mmd.cpu(), mm) and torch.equal(mvd.cpu(), mv)


@pytest.mark.parametrize('cfg', [(2, 16, 10, 16), (1, 32, 30, 16), (3, 16, 50, 16), (1, 4, 1, 3)])
def test_softargmax_gauss_fwd_bwd(cfg):
  B, S, K, Sg = cfg
  g = torch.Generator().manual_seed(S + K)
  heat = torch.randn(B, S, S, K, generator=g) * 2
  hd = heat.double().requires_grad_(True)
  gy, py = O.get_coord(hd, 2, S)
  gx, px = O.get_coord(hd, 1, S)
  mu = torch.stack([gy, gx], dim=2)
  maps = O.get_gaussian_maps(mu, [Sg, Sg], 10.0, 'rot')
  gm = torch.randn(maps.shape, generator=g)
  maps.backward(gm.double())
  dev, Cs, off = 'cuda', 64, 7
  mu_d, py_d, px_d = torch.empty(B, K, 2, device=dev), torch.empty(B, S, K, device=dev), torch.empty(B, S, K, device=dev)
  mh, ml = torch.zeros(B, Sg, Sg, Cs, device=dev), torch.zeros(B, Sg, Sg, Cs, device=dev)
  call('immb_softargmax_gauss_fwd', heat.to(dev), B, S, K, K, 10.0, mu_d, py_d, px_d, Sg, mh, ml, Cs, off, ST())
  assert float((mu_d.cpu().double() - mu).abs().max()) < 1e-5
  assert rel_err(py_d, py) < 1e-5 and rel_err(px_d, px) < 1e-5
  assert rel_err((mh + ml)[..., off:off + K], maps) < 1e-4
  assert float((mh + ml)[..., :off].abs().max()) == 0.0
  gfull = torch.zeros(B, Sg, Sg, Cs)
  gfull[..., off:off + K] = gm
  gh = torch.empty(B, S, S, K, device=dev)
  call('immb_softargmax_gauss_bwd', gfull.to(dev), Cs, off, mu_d, py_d, px_d, B, S, K, Sg, 10.0, gh, K, ST())
  assert rel_err(gh, hd.grad) < 1e-4
  full = torch.empty(B, Sg, Sg, K, device=dev)
  call('immb_gaussian_maps', mu_d, B, K, Sg, 10.0, full, ST())
  assert rel_err(full, maps) < 1e-4


def test_maxpool_fwd_bwd_first_max_wins():
  g = torch.Generator().manual_seed(3)
  x = torch.relu(torch.randn(2, 8, 6, 5, generator=g))       # many exact-zero ties, like post-ReLU VGG maps
  xd = x.double().requires_grad_(True)
  ref = O.max_pool_2x2(xd)
  go = torch.randn(ref.shape, generator=g)
  ref.backward(go.double())
  dev = 'cuda'
  xh, xl = split(x)
  oh, ol = torch.empty(2, 4, 3, 5, device=dev), torch.empty(2, 4, 3, 5, device=dev)
  call('immb_maxpool2x2_fwd', xh.to(dev), xl.to(dev), 2, 8, 6, 5, oh, ol, ST())
  assert rel_err(oh + ol, ref) < 1e-6
  gi = torch.empty(2, 8, 6, 5, device=dev)
  call('immb_maxpool2x2_bwd', go.to(dev), xh.to(dev), xl.to(dev), 2, 8, 6, 5, gi, ST())
  # positions with a strictly positive max are unambiguous; tie windows (all zeros) route to the first element
  pos = (ref.detach() > 0).float()
  up = lambda t: t.repeat_interleave(2, 1).repeat_interleave(2, 2)
  assert rel_err(gi.cpu() * up(pos), xd.grad.float() * up(pos)) < 1e-6
  assert rel_err(gi.cpu().view(2, 4, 2, 3, 2, 5).sum((2, 4)), go) < 1e-6      # each window passes its gradient once


def test_perceptual_sum_finalize_and_grads():
  B, R, h, C = 2, 16, 4, 6
  g = torch.Generator().manual_seed(9)
  f = torch.relu(torch.randn(2 * B, h, h, C, generator=g))
  mask = torch.rand(B, R, R, 1, generator=g)
  fd = f.double()
  fp = fd[B:].clone().requires_grad_(True)
  m = O.resize_bilinear(mask.double(), [h, h])
  l = (fd[:B] - fp) ** 2
  s = (l * m).mean()
  a = 2.3
  wl = a + 0.01 * (s - a)
  L = (l / wl * m).mean()
  (1000.0 * L).backward()
  dev = 'cuda'
  fh, fl = split(f)
  fh, fl = fh.to(dev), fl.to(dev)
  acc = torch.zeros(1, dtype=torch.float64, device=dev)
  call('immb_perceptual_level_sum', fh[:B], fl[:B], C, fh[B:], fl[B:], C, B, h, h, C, mask.to(dev), R, acc, ST())
  cnt = float(B * h * h * C)
  assert abs(float(acc.item()) / cnt - float(s)) / float(s) < 1e-5
  agg = torch.tensor([a], device=dev)
  lev, rec, coef = torch.empty(1, device=dev), torch.empty(1, device=dev), torch.empty(1, device=dev)
  call('immb_perceptual_finalize', acc, torch.tensor([cnt], dtype=torch.float64, device=dev), 1, agg, 1, lev, rec, coef, ST())
  assert abs(float(lev.item()) - float(L)) / float(L) < 1e-5
  assert abs(float(rec.item()) - 1000 * float(L)) / (1000 * float(L)) < 1e-5
  assert abs(float(agg.item()) - float(wl)) / float(wl) < 1e-6
  dyh, dyl = torch.empty(B, h, h, C, device=dev), torch.empty(B, h, h, C, device=dev)
  call('immb_vgg_bwd_combine', None, fh[:B], fl[:B], fh[B:], fl[B:], B, h, h, C, mask.to(dev), R, coef, dyh, dyl, ST())
  ref = fp.grad * (fp.detach() > 0)
  assert rel_err(dyh + dyl, ref) < 1e-4


def test_vgg_prologue_and_pred_grad():
  B, R = 2, 8
  g = torch.Generator().manual_seed(4)
  gt = torch.rand(B, R, R, 3, generator=g) * 255
  pred9 = torch.randn(B, R, R, 9, generator=g) * 50
  dev = 'cuda'
  oh, ol = torch.empty(2 * B, R, R, 1, device=dev), torch.empty(2 * B, R, R, 1, device=dev)
  call('immb_vgg_prologue', gt.to(dev), pred9.to(dev), 9, B, R, 0, oh, ol, ST())
  ims = torch.cat([gt, pred9[..., :3]], 0).double()
  ref = ims.mean(3, keepdim=True) / 255.0 - O.VGG_MEAN / 255.0
  assert rel_err(oh + ol, ref) < 1e-5
  mask = torch.rand(B, R, R, 1, generator=g)
  coef = torch.tensor([-0.37])
  gv = torch.randn(B, R, R, 1, generator=g)
  gh, gl = torch.empty(B, R, R, 9, device=dev), torch.empty(B, R, R, 9, device=dev)
  call('immb_pred_grad', gt.to(dev), pred9.to(dev), 9, mask.to(dev), coef.to(dev), gv.to(dev), 0, B, R, gh, gl, ST())
  ref = torch.zeros(B, R, R, 9, dtype=torch.float64)
  ref[..., :3] = -0.37 * mask.double() * (gt.double() - pred9[..., :3].double()) + gv.double() / (3 * 255.0)
  assert rel_err(gh + gl, ref) < 1e-5


def test_vgg_patch_prologue_and_adjoint():
  """conv1_1 (Cin=1, 3x3) == 1x1 conv over 3x3 patches; pred_grad applies the adjoint of the patch extraction."""
  B, R = 2, 16
  g = torch.Generator().manual_seed(6)
  gt = torch.rand(B, R, R, 3, generator=g) * 255
  pred12 = torch.zeros(B, R, R, 12)
  pred12[..., :9] = torch.randn(B, R, R, 9, generator=g) * 50
  dev = 'cuda'
  ph, pl = torch.empty(2 * B, R, R, 12, device=dev), torch.empty(2 * B, R, R, 12, device=dev)
  call('immb_vgg_prologue', gt.to(dev), pred12.to(dev), 12, B, R, 1, ph, pl, ST())
  ims = torch.cat([gt, pred12[..., :3]], 0).double().requires_grad_(True)
  gray = ims.mean(3, keepdim=True) / 255.0 - O.VGG_MEAN / 255.0
  pad = torch.nn.functional.pad(gray[..., 0], (1, 1, 1, 1))
  patches = torch.stack([pad[:, r:r + R, s:s + R] for r in range(3) for s in range(3)], dim=-1)
  assert rel_err((ph + pl)[..., :9], patches) < 1e-5
  assert float((ph + pl)[..., 9:].abs().max()) == 0.0
  gp = torch.randn(2 * B, R, R, 9, generator=g)
  patches.backward(gp.double())
  gfull = torch.zeros(B, R, R, 12)
  gfull[..., :9] = gp[B:]
  coef = torch.tensor([0.0])
  gh, gl = torch.empty(B, R, R, 12, device=dev), torch.empty(B, R, R, 12, device=dev)
  call('immb_pred_grad', gt.to(dev), pred12.to(dev), 12, None, coef.to(dev), gfull.to(dev), 1, B, R, gh, gl, ST())
  assert rel_err((gh + gl)[..., :3], ims.grad[B:]) < 1e-5


def test_vgg_conv1_1_fused():
  B, R = 2, 16
  g = torch.Generator().manual_seed(12)
  gt = torch.rand(B, R, R, 3, generator=g) * 255
  pred12 = torch.randn(B, R, R, 12, generator=g) * 60
  w = torch.randn(3, 3, 1, 64, generator=g)
  b = torch.randn(64, generator=g) * 0.1
  ims = torch.cat([gt, pred12[..., :3]], 0).double()
  gray = ims.mean(3, keepdim=True) / 255.0 - O.VGG_MEAN / 255.0
  ref = torch.relu(O.conv2d_same(gray, w.double(), b.double(), 1))
  dev = 'cuda'
  oh, ol = torch.empty(2 * B, R, R, 64, device=dev), torch.empty(2 * B, R, R, 64, device=dev)
  call('immb_vgg_conv1_1_fused', gt.to(dev), pred12.to(dev), 12, B, R, w.to(dev), b.to(dev), 64, oh, ol, 0, ST())
  assert rel_err(oh + ol, ref) < 1e-5
  # the two halves separately (the gt half runs on its own stream in the engine) write the same bits
  oh2, ol2 = torch.full_like(oh, float('nan')), torch.full_like(ol, float('nan'))
  call('immb_vgg_conv1_1_fused', gt.to(dev), None, 12, B, R, w.to(dev), b.to(dev), 64, oh2, ol2, 1, ST())
  assert torch.equal(oh2[:B], oh[:B]) and bool(torch.isnan(oh2[B:]).all())
  call('immb_vgg_conv1_1_fused', None, pred12.to(dev), 12, B, R, w.to(dev), b.to(dev), 64, oh2, ol2, 2, ST())
  assert torch.equal(oh2, oh) and torch.equal(ol2, ol)


def test_first_layer_rowwin_tcgen05():
  """7x7 / Cin=3 encoder conv_1 on the tensor cores via the staged row-window image (fwd + wgrad)."""
  N, R, Cout = 2, 32, 32
  g = torch.Generator().manual_seed(8)
  img = torch.rand(N, R, R, 3, generator=g) * 255
  w = torch.randn(7, 7, 3, Cout, generator=g) * 0.01
  b = torch.randn(Cout, generator=g)
  xd, wd = img.double().requires_grad_(True), w.double().requires_grad_(True)
  y_ref = O.conv2d_same(xd, wd, b.double(), 1)
  gy = torch.randn(y_ref.shape, generator=g)
  y_ref.backward(gy.double())
  dev = 'cuda'
  d = conv_desc(N, R, R, 3, Cout, 7, 1, xcs=4, engine=_lib.ENGINE_TC)
  d.x_layout = _lib.XLAYOUT_ROWWIN4
  assert _lib.lib().immb_conv_engine_for(d, 0) == _lib.ENGINE_TC and _lib.lib().immb_conv_engine_for(d, 2) == _lib.ENGINE_TC
  sh, sl = torch.empty(N, R, R + 8, 4, device=dev), torch.empty(N, R, R + 8, 4, device=dev)
  call('immb_stage_image_rowwin', img.to(dev), N, R, R, sh, sl, ST())
  ref = torch.zeros(N, R, R + 8, 4)
  ref[:, :, 3:3 + R, :3] = img
  assert rel_err(sh + sl, ref) < 1e-6
  wph, wpl = torch.empty(7, Cout, 32, device=dev), torch.empty(7, Cout, 32, device=dev)
  call('immb_pack_weights_rowwin', w.to(dev), Cout, wph, wpl, ST())
  y = torch.full((N, R, R, Cout), float('nan'), device=dev)
  call('immb_conv2d_fwd', d, sh, sl, None, wph, wpl, b.to(dev), y, None, ST())
  torch.cuda.synchronize()
  assert rel_err(y, y_ref) < 2e-5, rel_err(y, y_ref)
  gh, gl = split(gy)
  dw = torch.full((7, 7, 3, Cout), float('nan'), device=dev)
  ws = torch.empty(16, dtype=torch.uint8, device=dev)
  call('immb_conv2d_wgrad', d, sh, sl, gh.to(dev), gl.to(dev), dw, ws, 16, ST())
  torch.cuda.synchronize()
  assert rel_err(dw, wd.grad) < 2e-5, rel_err(dw, wd.grad)


def test_resize_align_corners_fwd_bwd():
  N, H, C, Ho = 2, 32, 8, 16
  x = torch.randn(N, H, H, C)
  xd = x.double().requires_grad_(True)
  ref = O.resize_bilinear(xd, [Ho, Ho], align_corners=True)
  go = torch.randn(ref.shape)
  ref.backward(go.double())
  dev = 'cuda'
  xh, xl = split(x)
  oh, ol = torch.empty(N, Ho, Ho, C, device=dev), torch.empty(N, Ho, Ho, C, device=dev)
  call('immb_resize_ac_fwd', xh.to(dev), xl.to(dev), C, N, H, H, C, Ho, Ho, oh, ol, C, ST())
  assert rel_err(oh + ol, ref) < 1e-5
  gi = torch.empty(N, H, H, C, device=dev)
  call('immb_resize_ac_bwd', go.to(dev), C, N, H, H, C, Ho, Ho, gi, ST())
  assert rel_err(gi, xd.grad) < 1e-5


def test_clip_adam_multi_tensor():
  """Two tensors in one flat buffer: one gets clipped (norm>1), one does not; wd only on the first."""
  dev = 'cuda'
  g = torch.Generator().manual_seed(11)
  n0, n1 = 3000, 37
  p = torch.randn(n0 + n1, generator=g)
  gr = torch.cat([torch.randn(n0, generator=g), torch.randn(n1, generator=g) * 1e-3])
  m0, v0 = torch.rand(n0 + n1, generator=g) * 0.01, torch.rand(n0 + n1, generator=g) * 1e-4
  wd = [1e-5, 0.0]
  ct, co, cl = [], [], []
  for t, (o, n) in enumerate([(0, n0), (n0, n1)]):
    for c0 in range(0, n, 2048):
      ct.append(t); co.append(o + c0); cl.append(min(2048, n - c0))
  T = lambda a, dt: torch.tensor(a, dtype=dt, device=dev)
  pd, gd, md, vd = p.to(dev), gr.to(dev), m0.to(dev), v0.to(dev)
  sq = torch.zeros(4, dtype=torch.float64, device=dev)
  world = 2.0
  args = (T(ct, torch.int32), T(co, torch.int64), T(cl, torch.int32), len(ct), T(wd, torch.float32), 1.0 / world)
  call('immb_adam_norms', pd, gd, n0 + n1, *args, sq, sq[2:], ST())
  lr, t = 1e-3, 3
  lr_t = lr * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
  call('immb_adam_apply', pd, gd, md, vd, n0 + n1, *args, sq, 1.0, lr_t, 0.9, 0.999, 1e-8, ST())
  for (o, n), w in zip([(0, n0), (n0, n1)], wd):
    pe, ge = p[o:o + n].double(), gr[o:o + n].double() / world
    ge = ge + w * pe
    gc = O.clip_by_norm(ge, 1.0)
    var, mm, vv = O.adam_step(pe, gc, m0[o:o + n].double(), v0[o:o + n].double(), lr, t)
    assert rel_err(pd[o:o + n], var) < 1e-6
    assert rel_err(md[o:o + n], mm) < 1e-5 and rel_err(vd[o:o + n], vv) < 1e-5
  wl, tot = torch.empty(1, device=dev), torch.empty(1, device=dev)
  call('immb_total_loss', T([5.0], torch.float32), sq[2:], T(wd, torch.float32), 2, wl, tot, ST())
  ref_wl = 0.5 * 1e-5 * float((p[:n0].double() ** 2).sum())
  assert abs(float(wl.item()) - ref_wl) / ref_wl < 1e-5 and abs(float(tot.item()) - 5.0 - ref_wl) < 1e-5


def test_error_reporting_no_throw():
  d = conv_desc(1, 8, 8, 4, 4, 3, 1)
  d.kh = 9
  with pytest.raises(_lib.ImmbError) as e:
    call('immb_conv2d_fwd', d, torch.zeros(1, device='cuda'), None, torch.zeros(1, device='cuda'), None, None, None,
         torch.zeros(1, device='cuda'), None, ST())
  assert 'kernel size' in str(e.value)
  with pytest.raises(_lib.ImmbError):
    call('immb_split_planes', torch.zeros(4), torch.zeros(4), None, 4, ST())       # CPU tensors are refused


@pytest.mark.parametrize('case', [(2, 32, 32, 32, 32, 3), (1, 16, 16, 288, 256, 3), (3, 16, 16, 64, 96, 3), (2, 32, 32, 3, 32, 7)])
def test_conv_fwd_fused_bn_statistics(case):
  """immb_conv2d_fwd_bnstats: the conv output is bit-identical to immb_conv2d_fwd and the per-channel sum / sum of
  squares accumulated in the epilogue match the oracle's batch moments of that output (nn_utils.py:201)."""
  N, H, W, Cin, Cout, k = case
  g = torch.Generator().manual_seed(sum(case))
  dev = 'cuda'
  first = k == 7
  w = torch.randn(k, k, Cin, Cout, generator=g) * 0.05
  b = torch.randn(Cout, generator=g)
  if first:
    img = torch.rand(N, H, W, 3, generator=g) * 255
    d = conv_desc(N, H, W, 3, Cout, 7, 1, xcs=4, engine=_lib.ENGINE_TC)
    d.x_layout = _lib.XLAYOUT_ROWWIN4
    xh, xl = torch.empty(N, H, W + 8, 4, device=dev), torch.empty(N, H, W + 8, 4, device=dev)
    call('immb_stage_image_rowwin', img.to(dev), N, H, W, xh, xl, ST())
    wph, wpl = torch.empty(7, Cout, 32, device=dev), torch.empty(7, Cout, 32, device=dev)
    call('immb_pack_weights_rowwin', w.to(dev), Cout, wph, wpl, ST())
  else:
    x = torch.randn(N, H, W, Cin, generator=g)
    d = conv_desc(N, H, W, Cin, Cout, 3, 1, None, engine=_lib.ENGINE_TC)
    xh, xl = (t.to(dev) for t in split(x))
    cp = d.cin_pad
    wph, wpl = torch.empty(9, Cout, cp, device=dev), torch.empty(9, Cout, cp, device=dev)
    whh, whl = torch.empty(9, cp, Cout, device=dev), torch.empty(9, cp, Cout, device=dev)
    call('immb_pack_weights', w.to(dev), 3, 3, Cin, Cout, cp, Cout, wph, wpl, whh, whl, ST())
  rows = int(_lib.lib().immb_conv2d_fwd_stats_rows(d))
  assert rows > 0 and rows % 8 == 0
  y0 = torch.empty(N, H, W, Cout, device=dev)
  call('immb_conv2d_fwd', d, xh, xl, None, wph, wpl, b.to(dev), y0, None, ST())
  y1 = torch.full_like(y0, float('nan'))
  part = torch.full((rows * 2 * Cout,), float('nan'), dtype=torch.float64, device=dev)
  call('immb_conv2d_fwd_bnstats', d, xh, xl, wph, wpl, b.to(dev), y1, part, part.numel(), ST())
  sums = torch.zeros(2 * Cout, dtype=torch.float64, device=dev)
  call('immb_bn_stats_from_partials', part, rows, Cout, sums, ST())
  torch.cuda.synchronize()
  assert torch.equal(y0, y1)
  yd = y0.double().reshape(-1, Cout)
  ref = torch.cat([yd.sum(0), (yd * yd).sum(0)])
  assert bool(torch.isfinite(part).all())
  np.testing.assert_allclose(sums.cpu().numpy(), ref.cpu().numpy(), rtol=2e-6, atol=1e-6 * float(ref.abs().max()))
  # single-pass TF32 / shapes outside the pair kernel: not offered
  d1 = conv_desc(N, 8, 8, 32, 32, 3, 1, None, engine=_lib.ENGINE_TC)
  assert int(_lib.lib().immb_conv2d_fwd_stats_rows(d1)) == 0


@pytest.mark.parametrize('case', [(2, 32, 32, 32, 32, 1), (1, 16, 16, 256, 256, 1), (3, 16, 16, 64, 128, 0)])
def test_conv_dgrad_fused_bn_backward_sums(case):
  """immb_conv2d_dgrad_bnreduce: dx is bit-identical to immb_conv2d_dgrad, and the per-channel sums of the producing
  layer's BN backward (sum dz, sum dz*xhat; nn_utils.py:201-209) accumulated in the epilogue match immb_bn_bwd_reduce
  run on that dx."""
  N, H, W, Cin, Cout, relu = case
  g = torch.Generator().manual_seed(sum(case) + 5)
  dev = 'cuda'
  w = torch.randn(3, 3, Cin, Cout, generator=g) * 0.05
  gy = torch.randn(N, H, W, Cout, generator=g)
  y_prev = torch.randn(N, H, W, Cin, generator=g) * 3 + 1          # raw conv output of the producing layer
  scale, shift = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g)
  mean, invstd = torch.randn(Cin, generator=g), torch.rand(Cin, generator=g) + 0.2
  d = conv_desc(N, H, W, Cin, Cout, 3, 1, None, engine=_lib.ENGINE_TC)
  rows = int(_lib.lib().immb_conv2d_dgrad_stats_rows(d))
  assert rows > 0
  cp = d.cin_pad
  wph, wpl = torch.empty(9, Cout, cp, device=dev), torch.empty(9, Cout, cp, device=dev)
  whh, whl = torch.empty(9, cp, Cout, device=dev), torch.empty(9, cp, Cout, device=dev)
  call('immb_pack_weights', w.to(dev), 3, 3, Cin, Cout, cp, Cout, wph, wpl, whh, whl, ST())
  gh, gl = (t.to(dev) for t in split(gy))
  dx0 = torch.empty(N, H, W, Cin, device=dev)
  call('immb_conv2d_dgrad', d, gh, gl, None, whh, whl, dx0, ST())
  dx1 = torch.full_like(dx0, float('nan'))
  part = torch.full((rows * 2 * Cin,), float('nan'), dtype=torch.float64, device=dev)
  args = [t.to(dev) for t in (y_prev, scale, shift, mean, invstd)]
  call('immb_conv2d_dgrad_bnreduce', d, gh, gl, whh, whl, dx1, args[0], Cin, args[1], args[2], args[3], args[4], relu,
       part, part.numel(), ST())
  sums = torch.zeros(2 * Cin, dtype=torch.float64, device=dev)
  call('immb_bn_stats_from_partials', part, rows, Cin, sums, ST())
  ref = torch.zeros(2 * Cin, dtype=torch.float64, device=dev)
  call('immb_bn_bwd_reduce', dx0, Cin, args[0], Cin, N * H * W, Cin, args[1], args[2], args[3], args[4], relu, ref,
       None, 0, ST())
  torch.cuda.synchronize()
  assert torch.equal(dx0, dx1) and bool(torch.isfinite(part).all())
  dz = dx0.double().cpu()
  if relu:
    dz = dz * ((y_prev.double() * scale.double() + shift.double()) > 0)
  xh = (y_prev.double() - mean.double()) * invstd.double()
  want = torch.cat([dz.reshape(-1, Cin).sum(0), (dz * xh).reshape(-1, Cin).sum(0)])
  tol = 1e-5 * float(want.abs().max())
  np.testing.assert_allclose(sums.cpu().numpy(), want.numpy(), rtol=1e-4, atol=tol)
  np.testing.assert_allclose(sums.cpu().numpy(), ref.cpu().numpy(), rtol=1e-4, atol=tol)


@pytest.mark.parametrize('use_mask', [True, False])
def test_maxpool_fused_level_sum(use_mask):
  """immb_maxpool2x2_fwd_levelsum = immb_maxpool2x2_fwd on [gt ; pred] + immb_perceptual_level_sum of the level
  (imm_model.py:143-147 with _loss_mask's subsampled mask) in one pass."""
  B, H, C, R = 3, 16, 8, 64
  g = torch.Generator().manual_seed(31 + use_mask)
  x = torch.relu(torch.randn(2 * B, H, H, C, generator=g))
  mask = torch.rand(B, R, R, 1, generator=g)
  dev = 'cuda'
  xh, xl = (t.to(dev) for t in split(x))
  md = mask.to(dev) if use_mask else None
  o0h, o0l = torch.empty(2 * B, H // 2, H // 2, C, device=dev), torch.empty(2 * B, H // 2, H // 2, C, device=dev)
  call('immb_maxpool2x2_fwd', xh, xl, 2 * B, H, H, C, o0h, o0l, ST())
  acc0 = torch.zeros(1, dtype=torch.float64, device=dev)
  call('immb_perceptual_level_sum', xh[:B], xl[:B], C, xh[B:], xl[B:], C, B, H, H, C, md, R, acc0, ST())
  o1h, o1l = torch.full_like(o0h, float('nan')), torch.full_like(o0l, float('nan'))
  acc1 = torch.zeros(1, dtype=torch.float64, device=dev)
  call('immb_maxpool2x2_fwd_levelsum', xh, xl, B, H, H, C, o1h, o1l, md, R, acc1, ST())
  torch.cuda.synchronize()
  assert torch.equal(o0h, o1h) and torch.equal(o0l, o1l)
  assert abs(float(acc0) - float(acc1)) <= 1e-12 * abs(float(acc0))
  xs = (xh + xl).double().cpu()
  m = mask[:, ::R // H, ::R // H].double() if use_mask else 1.0
  want = float((m * (xs[:B] - xs[B:]) ** 2).sum())
  assert abs(float(acc1) - want) <= 1e-6 * want


@pytest.mark.parametrize('with_loss', [True, False])
def test_maxpool_bwd_fused_combine(with_loss):
  """immb_maxpool2x2_bwd_combine = immb_maxpool2x2_bwd followed by immb_vgg_bwd_combine (bit-identical planes)."""
  B, H, C, R = 2, 16, 8, 32
  g = torch.Generator().manual_seed(41 + with_loss)
  f = torch.relu(torch.randn(2 * B, H, H, C, generator=g))
  f[B:, ::2, ::2] = f[B:, 1::2, ::2]                       # ties inside pooling windows: first-max-wins must hold
  g_out = torch.randn(B, H // 2, H // 2, C, generator=g)
  mask = torch.rand(B, R, R, 1, generator=g)
  coef = torch.tensor([0.37])
  dev = 'cuda'
  fh, fl = (t.to(dev) for t in split(f))
  cd = coef.to(dev) if with_loss else None
  g_in = torch.empty(B, H, H, C, device=dev)
  call('immb_maxpool2x2_bwd', g_out.to(dev), fh[B:], fl[B:], B, H, H, C, g_in, ST())
  d0h, d0l = torch.empty(B, H, H, C, device=dev), torch.empty(B, H, H, C, device=dev)
  call('immb_vgg_bwd_combine', g_in, fh[:B], fl[:B], fh[B:], fl[B:], B, H, H, C, mask.to(dev), R, cd, d0h, d0l, ST())
  d1h, d1l = torch.full_like(d0h, float('nan')), torch.full_like(d0l, float('nan'))
  call('immb_maxpool2x2_bwd_combine', g_out.to(dev), fh[:B], fl[:B], fh[B:], fl[B:], B, H, H, C, mask.to(dev), R, cd,
       d1h, d1l, ST())
  torch.cuda.synchronize()
  assert torch.equal(d0h, d1h) and torch.equal(d0l, d1l)


def test_vgg_conv1_1_backward_fused():
  """immb_vgg_conv1_1_bwd_fused (conv1_1 dgrad + gray/normalise adjoint + 'input' level term in one exact-fp32 kernel)
  vs autograd through the oracle's prologue + conv1_1, and vs the two-kernel path it replaces."""
  B, R, pcs = 2, 32, 12
  g = torch.Generator().manual_seed(51)
  gt = torch.rand(B, R, R, 3, generator=g) * 255
  pred12 = torch.randn(B, R, R, pcs, generator=g) * 60
  w = torch.randn(3, 3, 1, 64, generator=g)
  dy = torch.randn(B, R, R, 64, generator=g)
  mask = torch.rand(B, R, R, 1, generator=g)
  coef = torch.tensor([0.013])
  dev = 'cuda'
  pd = pred12[..., :3].double().clone().requires_grad_(True)
  gray = pd.mean(3, keepdim=True) / 255.0 - O.VGG_MEAN / 255.0
  y = O.conv2d_same(gray, w.double(), None, 1)
  y.backward(dy.double())
  want = pd.grad + float(coef) * mask.double() * (gt.double() - pred12[..., :3].double())
  dh, dl = (t.to(dev) for t in split(dy))
  gh, gl = torch.full((B, R, R, pcs), float('nan'), device=dev), torch.full((B, R, R, pcs), float('nan'), device=dev)
  call('immb_vgg_conv1_1_bwd_fused', dh, dl, w.to(dev), 64, gt.to(dev), pred12.to(dev), pcs, mask.to(dev), coef.to(dev),
       B, R, gh, gl, ST())
  torch.cuda.synchronize()
  got = (gh + gl).cpu()
  assert rel_err(got[..., :3], want) < 1e-5
  assert float(got[..., 3:].abs().max()) == 0.0
  # the path it replaces: 1x1 dgrad over the patch tensor on the tensor cores + pred_grad
  d = conv_desc(B, R, R, 9, 64, 1, 1, xcs=12, engine=_lib.ENGINE_TC)
  wp_h, wp_l = torch.empty(1, 64, 32, device=dev), torch.empty(1, 64, 32, device=dev)
  wh_h, wh_l = torch.empty(1, 32, 64, device=dev), torch.empty(1, 32, 64, device=dev)
  call('immb_pack_weights', w.reshape(1, 1, 9, 64).contiguous().to(dev), 1, 1, 9, 64, 32, 64, wp_h, wp_l, wh_h, wh_l, ST())
  dx = torch.zeros(B, R, R, 12, device=dev)
  call('immb_conv2d_dgrad', d, dh, dl, None, wh_h, wh_l, dx, ST())
  g2h, g2l = torch.empty_like(gh), torch.empty_like(gl)
  call('immb_pred_grad', gt.to(dev), pred12.to(dev), pcs, mask.to(dev), coef.to(dev), dx, 1, B, R, g2h, g2l, ST())
  assert rel_err(got, (g2h + g2l).cpu()) < 2e-5
