"""Scaled-fp16 split operands ("H16 planes", include/imm_b200.h): kernel-level parity of the kind::f16 convolutions and
of the glue kernels that read / write fp16 planes, against the fp64 oracle.  The bars are the 3xTF32 ones: the format
carries the same 22 significant bits per operand."""
import numpy as np
import pytest
import torch

from imm_b200 import _lib
from imm_b200._lib import call
from oracle import imm_oracle as O
from tests.gpu_util import rel_err
from tests.test_gpu_ops import conv_desc

pytestmark = pytest.mark.gpu
ST = lambda: _lib.stream_ptr()
DEV = 'cuda'


def h16_exp(amax, target=12):
  if amax <= 0:
    return 0
  m, ex = np.frexp(np.float32(amax))
  return int(target - ex)


def h16_split(t, e=None):
  """host-side restatement of the format: (hi, lo fp16 CUDA planes, scale record int32[2])."""
  t = t.float()
  if e is None:
    e = h16_exp(float(t.abs().max()))
  s = t * (2.0 ** e)
  hi = s.half()
  lo = ((s - hi.float()) * 2048.0).half()
  rec = torch.tensor([e, 0], dtype=torch.int32, device=DEV)
  return hi.to(DEV).contiguous(), lo.to(DEV).contiguous(), rec


def h16_join(hi, lo, rec):
  e = int(rec[0].item())
  return (hi.float() + lo.float() / 2048.0).cpu().double() * (2.0 ** -e)


def test_h16_format_round_trip_and_split_kernel():
  g = torch.Generator().manual_seed(0)
  v = torch.randn(4, 8, 8, 32, generator=g) * torch.logspace(-6, 1, 32)      # 7 decades inside one tensor
  hi, lo, rec = h16_split(v)
  assert rel_err(h16_join(hi, lo, rec), v) < 3e-7
  # the kernel writes the same planes and tracks the largest magnitude in the record
  khi, klo = torch.empty_like(hi), torch.empty_like(lo)
  call('immb_split_planes', v.to(DEV), khi, klo, v.numel(), rec, ST())
  torch.cuda.synchronize()
  assert torch.equal(khi, hi) and torch.equal(klo, lo)
  amax = rec[1:2].view(torch.float32).item()
  assert amax == float(v.abs().max())
  # immb_scale_update: next exponent from the observed maximum, maximum cleared, no overflow counted
  recs = torch.zeros((3, 2), dtype=torch.int32, device=DEV)
  recs[0] = rec
  recs[1, 0] = 5                                                   # untouched tensor keeps its exponent
  recs[2, 0] = 20
  recs[2, 1:2].view(torch.float32).fill_(1.0)                      # 1.0 * 2^20 does not fit fp16 -> overflow
  ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
  call('immb_scale_update', recs, 3, ovf, ST())
  torch.cuda.synchronize()
  assert recs[:, 1].tolist() == [0, 0, 0]
  # delayed-scaling tensors aim at 2^8 (256x headroom); weights (exact, same-step scaling) at 2^12
  assert recs[0, 0].item() == h16_exp(float(v.abs().max()), target=8) and recs[1, 0].item() == 5 and recs[2, 0].item() == 7
  assert ovf.item() == 1


H16_CASES = [
  # N, H, W, Cin, Cout, k, stride, xcs
  (2, 16, 16, 32, 32, 3, 1, None),       # encoder conv_2 shape: half-filled 64-channel chunks on both operands
  (1, 32, 32, 64, 128, 3, 1, None),
  (2, 16, 16, 266, 256, 3, 1, 288),      # renderer conv_1: ragged last K chunk (288 = 4 x 64 + 32), two N tiles
  (1, 128, 128, 32, 32, 3, 1, None),
  (1, 32, 32, 64, 32, 3, 1, None),
  (2, 16, 16, 512, 512, 3, 1, None),     # VGG conv4_x: K = 4608
  (1, 32, 32, 32, 9, 3, 1, None),        # renderer last conv: Cout = 9 in a 16-channel-stride tensor
  (3, 16, 16, 64, 96, 3, 1, None),       # N tile of 96, odd image count
  (1, 16, 16, 320, 192, 3, 1, None),
  (4, 8, 8, 128, 128, 3, 1, None),       # 8x8 maps: conv_tc_kernel forward / dgrad, halo wgrad
  (2, 32, 32, 32, 64, 3, 2, None),       # stride 2, C = 32: both column parities of the view in one 64-channel window
  (2, 64, 64, 64, 128, 3, 2, None),      # stride 2, C = 64: one window per column parity
  (1, 32, 32, 128, 256, 3, 2, None),     # stride 2, C = 128: two channel chunks, two N tiles
]


@pytest.mark.parametrize('case', H16_CASES)
def test_conv_h16_engine(case):
  N, H, W, Cin, Cout, k, stride, xcs = case
  g = torch.Generator().manual_seed(sum(case[:7]))
  xcs_ = xcs or Cin
  x = torch.randn(N, H, W, xcs_, generator=g) * 3.0
  x[..., Cin:] = 0
  w = torch.randn(k, k, Cin, Cout, generator=g) * 0.1
  b = torch.randn(Cout, generator=g)
  ycs = (Cout + 7) // 8 * 8
  d = conv_desc(N, H, W, Cin, Cout, k, stride, xcs, engine=_lib.ENGINE_TC, precision=_lib.PREC_F16X3, ycs=ycs)
  engines = [_lib.lib().immb_conv_engine_for(d, op) for op in range(3)]
  assert engines[:2] == [_lib.ENGINE_TC] * 2
  has_wgrad = engines[2] == _lib.ENGINE_TC
  assert has_wgrad == (k == 3 and ((stride == 1 and H % 4 == 0 and W % 8 == 0) or (stride == 2 and (H // 2) % 4 == 0 and (W // 2) % 8 == 0)))
  tol = 2e-5 if k * k * Cin < 2048 else 6e-5
  xd = x.double()[..., :Cin].clone().requires_grad_(True)
  wd = w.double().clone().requires_grad_(True)
  y_ref = O.conv2d_same(xd, wd, b.double(), stride)
  gy = torch.randn(y_ref.shape, generator=g) * 1e-5            # gradients live many binades below the activations
  y_ref.backward(gy.double())
  taps, cp = k * k, d.cin_pad
  f16 = dict(dtype=torch.float16, device=DEV)
  wp_h, wp_l = torch.empty(taps, Cout, cp, **f16), torch.empty(taps, Cout, cp, **f16)
  wh_h, wh_l = torch.empty(taps, cp, ycs, **f16), torch.empty(taps, cp, ycs, **f16)
  w_amax = torch.tensor([float(w.abs().max())], device=DEV)
  w_rec = torch.zeros(2, dtype=torch.int32, device=DEV)
  call('immb_pack_weights', w.to(DEV), k, k, Cin, Cout, cp, ycs, wp_h, wp_l, wh_h, wh_l, w_amax, w_rec, ST())
  xh, xl, x_rec = h16_split(x)
  d.x_scale, d.w_scale = x_rec.data_ptr(), w_rec.data_ptr()
  y = torch.full((N, d.Ho, d.Wo, ycs), float('nan'), device=DEV)
  call('immb_conv2d_fwd', d, xh, xl, None, wp_h, wp_l, b.to(DEV), y, None, ST())
  torch.cuda.synchronize()
  assert w_rec[0].item() == h16_exp(float(w.abs().max()))
  assert rel_err(y[..., :Cout], y_ref) < tol, ('fwd', rel_err(y[..., :Cout], y_ref))
  if ycs > Cout:
    assert float(y[..., Cout:].abs().max()) == 0.0
  gyp = torch.zeros(N, d.Ho, d.Wo, ycs)
  gyp[..., :Cout] = gy
  gh, gl, g_rec = h16_split(gyp)
  d.y_scale = g_rec.data_ptr()
  dx = torch.full((N, H, W, xcs_), float('nan'), device=DEV)
  call('immb_conv2d_dgrad', d, gh, gl, None, wh_h, wh_l, dx, ST())
  torch.cuda.synchronize()
  assert rel_err(dx[..., :Cin], xd.grad) < tol, ('dgrad', rel_err(dx[..., :Cin], xd.grad))
  if xcs_ > Cin:
    assert float(dx[..., Cin:].abs().max()) == 0.0
  if has_wgrad:
    dw = torch.full((k, k, Cin, Cout), float('nan'), device=DEV)
    ws = torch.empty(16, dtype=torch.uint8, device=DEV)
    call('immb_conv2d_wgrad', d, xh, xl, gh, gl, dw, ws, 16, ST())
    torch.cuda.synchronize()
    assert rel_err(dw, wd.grad) < tol, ('wgrad', rel_err(dw, wd.grad))
  # bias + ReLU epilogue writing H16 planes with a given exponent; the record receives the largest output
  d2 = conv_desc(N, H, W, Cin, Cout, k, stride, xcs, engine=_lib.ENGINE_TC, epilogue=_lib.EPI_BIAS_RELU,
                 precision=_lib.PREC_F16X3, ycs=ycs)
  y_rec = torch.tensor([h16_exp(float(torch.relu(y_ref).max())), 0], dtype=torch.int32, device=DEV)
  d2.x_scale, d2.w_scale, d2.y_scale = x_rec.data_ptr(), w_rec.data_ptr(), y_rec.data_ptr()
  yh, yl = torch.zeros(N, d.Ho, d.Wo, ycs, **f16), torch.zeros(N, d.Ho, d.Wo, ycs, **f16)
  call('immb_conv2d_fwd', d2, xh, xl, None, wp_h, wp_l, b.to(DEV), yh, yl, ST())
  torch.cuda.synchronize()
  assert rel_err(h16_join(yh, yl, y_rec)[..., :Cout], torch.relu(y_ref)) < tol
  amax = y_rec[1:2].view(torch.float32).item()
  assert abs(amax - float(torch.relu(y_ref).max())) <= 1e-4 * amax


def test_conv_h16_two_pass_frozen_weights_and_relu_backward():
  """IMMB_PREC_F16X2 (frozen tower): hi*w + lo*w with the weights rounded to their fp16 hi plane; the dgrad epilogue
  applies the backward of the ReLU that produced the conv's input and writes the previous layer's dy as H16 planes."""
  N, H, W, Cin, Cout = 2, 32, 32, 64, 128
  g = torch.Generator().manual_seed(7)
  x = torch.relu(torch.randn(N, H, W, Cin, generator=g))
  w = torch.randn(3, 3, Cin, Cout, generator=g) * 0.05
  b = torch.randn(Cout, generator=g) * 0.1
  d = conv_desc(N, H, W, Cin, Cout, 3, 1, None, engine=_lib.ENGINE_TC, precision=_lib.PREC_F16X2, ycs=Cout,
                epilogue=_lib.EPI_BIAS_RELU)
  f16 = dict(dtype=torch.float16, device=DEV)
  wp_h, wp_l = torch.empty(9, Cout, Cin, **f16), torch.empty(9, Cout, Cin, **f16)
  wh_h, wh_l = torch.empty(9, Cin, Cout, **f16), torch.empty(9, Cin, Cout, **f16)
  w_amax = torch.tensor([float(w.abs().max())], device=DEV)
  w_rec = torch.zeros(2, dtype=torch.int32, device=DEV)
  call('immb_pack_weights', w.to(DEV), 3, 3, Cin, Cout, Cin, Cout, wp_h, wp_l, wh_h, wh_l, w_amax, w_rec, ST())
  torch.cuda.synchronize()
  e_w = int(w_rec[0].item())
  w_used = (wh_h.float().cpu().double() * 2.0 ** -e_w).view(3, 3, Cin, Cout)          # the weights the product sees
  assert rel_err(w_used, w) < 6e-4
  xd = x.double().clone().requires_grad_(True)
  y_ref = torch.relu(O.conv2d_same(xd, w_used, b.double(), 1))
  gy = torch.randn(y_ref.shape, generator=g) * 1e-4
  y_ref.backward(gy.double())
  xh, xl, x_rec = h16_split(x)
  y_rec = torch.tensor([h16_exp(float(y_ref.max())), 0], dtype=torch.int32, device=DEV)
  d.x_scale, d.w_scale, d.y_scale = x_rec.data_ptr(), w_rec.data_ptr(), y_rec.data_ptr()
  yh, yl = torch.zeros(N, H, W, Cout, **f16), torch.zeros(N, H, W, Cout, **f16)
  call('immb_conv2d_fwd', d, xh, xl, None, wp_h, wp_l, b.to(DEV), yh, yl, ST())
  torch.cuda.synchronize()
  assert rel_err(h16_join(yh, yl, y_rec), y_ref) < 2e-5
  # dy of this layer (ReLU backward applied by the caller), then dgrad * [x > 0] -> previous layer's dy planes
  dy = gy.double() * (y_ref > 0)
  gh, gl, g_rec = h16_split(dy.float())
  want = xd.grad * (x > 0)
  o_rec = torch.tensor([h16_exp(float(want.abs().max())), 0], dtype=torch.int32, device=DEV)
  d.y_scale = g_rec.data_ptr()
  assert _lib.lib().immb_conv2d_dgrad_relu_supported(d)
  oh, ol = torch.zeros(N, H, W, Cin, **f16), torch.zeros(N, H, W, Cin, **f16)
  call('immb_conv2d_dgrad_relu', d, gh, gl, wh_h, wh_l, xh, Cin, oh, ol, o_rec, ST())
  torch.cuda.synchronize()
  assert rel_err(h16_join(oh, ol, o_rec), want) < 2e-5


def test_h16_glue_kernels_match_their_fp32_plane_twins():
  """bn_apply (+x2 upsample), max-pool (+level sum), perceptual level sum, BN-backward apply and bias gradient give the
  same numbers through fp16 planes as through fp32 planes (to the 2^-22 of the formats)."""
  g = torch.Generator().manual_seed(3)
  N, H, W, C = 2, 16, 16, 64
  y = torch.randn(N, H, W, C, generator=g).to(DEV) * 4
  scale, shift = (torch.rand(C, generator=g) + 0.5).to(DEV), torch.randn(C, generator=g).to(DEV)
  f16 = dict(dtype=torch.float16, device=DEV)
  for up in (0, 1):
    s = 2 if up else 1
    rh, rl = torch.empty(N, H * s, W * s, C, device=DEV), torch.empty(N, H * s, W * s, C, device=DEV)
    call('immb_bn_apply', y, N, H, W, C, C, scale, shift, 1, up, rh, rl, C, ST())
    ref = (rh + rl).cpu().double()
    rec = torch.tensor([h16_exp(float(ref.abs().max())), 0], dtype=torch.int32, device=DEV)
    oh, ol = torch.empty(N, H * s, W * s, C, **f16), torch.empty(N, H * s, W * s, C, **f16)
    call('immb_bn_apply', y, N, H, W, C, C, scale, shift, 1, up, oh, ol, C, rec, ST())
    torch.cuda.synchronize()
    assert rel_err(h16_join(oh, ol, rec), ref) < 5e-7
    assert abs(rec[1:2].view(torch.float32).item() - float(ref.abs().max())) <= 1e-6 * float(ref.abs().max())
  # pool + level sum on a [gt ; pred] stack
  B = 2
  act = torch.relu(torch.randn(2 * B, H, W, C, generator=g)) * 2
  mask = torch.rand(B, 32, 32, 1, generator=g).to(DEV)
  ah, al = O.round_tf32(act), O.round_tf32(act - O.round_tf32(act))
  ph, pl_ = torch.empty(2 * B, H // 2, W // 2, C, device=DEV), torch.empty(2 * B, H // 2, W // 2, C, device=DEV)
  acc32 = torch.zeros(1, dtype=torch.float64, device=DEV)
  call('immb_maxpool2x2_fwd_levelsum', ah.to(DEV), al.to(DEV), B, H, W, C, ph, pl_, mask, 32, acc32, ST())
  xh, xl, x_rec = h16_split(act)
  o_rec = torch.tensor([int(x_rec[0].item()), 0], dtype=torch.int32, device=DEV)
  qh, ql = torch.empty(2 * B, H // 2, W // 2, C, **f16), torch.empty(2 * B, H // 2, W // 2, C, **f16)
  acc16 = torch.zeros(1, dtype=torch.float64, device=DEV)
  call('immb_maxpool2x2_fwd_levelsum', xh, xl, B, H, W, C, qh, ql, mask, 32, acc16, x_rec, o_rec, ST())
  acc_ls = torch.zeros(1, dtype=torch.float64, device=DEV)
  call('immb_perceptual_level_sum', xh[:B], xl[:B], C, xh[B:], xl[B:], C, B, H, W, C, mask, 32, acc_ls, x_rec, ST())
  torch.cuda.synchronize()
  assert rel_err(h16_join(qh, ql, o_rec), (ph + pl_)) < 5e-7
  assert abs(acc16.item() - acc32.item()) <= 1e-6 * acc32.item()
  assert abs(acc_ls.item() - acc32.item()) <= 1e-6 * acc32.item()
  # BN backward apply + bias gradient
  gup = torch.randn(N, H, W, C, generator=g).to(DEV) * 1e-5
  mean, invstd = torch.randn(C, generator=g).to(DEV) * 0.1, (torch.rand(C, generator=g) + 0.5).to(DEV)
  sums = torch.randn(2 * C, generator=g).double().to(DEV) * 1e-4
  outs = []
  for fmt in ('f32', 'h16'):
    dg, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    dacc = torch.zeros(C, dtype=torch.float64, device=DEV)
    if fmt == 'f32':
      dh, dl = torch.empty(N, H, W, C, device=DEV), torch.empty(N, H, W, C, device=DEV)
      call('immb_bn_bwd_apply', gup, C, y, C, N * H * W, C, scale, shift, mean, invstd, 1, sums, dh, dl, dg, db, dacc, None, 0, ST())
      val = (dh + dl).cpu().double()
      bacc = torch.zeros(C, dtype=torch.float64, device=DEV)
      call('immb_bias_grad', dh, dl, C, N * H * W, C, bacc, ST())
    else:
      rec = torch.tensor([h16_exp(float(outs[0][0].abs().max())), 0], dtype=torch.int32, device=DEV)
      dh, dl = torch.empty(N, H, W, C, **f16), torch.empty(N, H, W, C, **f16)
      call('immb_bn_bwd_apply', gup, C, y, C, N * H * W, C, scale, shift, mean, invstd, 1, sums, dh, dl, dg, db, dacc, None, 0, rec, ST())
      torch.cuda.synchronize()
      val = h16_join(dh, dl, rec)
      bacc = torch.zeros(C, dtype=torch.float64, device=DEV)
      call('immb_bias_grad', dh, dl, C, N * H * W, C, bacc, rec, ST())
    torch.cuda.synchronize()
    outs.append((val, dacc.cpu(), bacc.cpu()))
  assert rel_err(outs[1][0], outs[0][0]) < 5e-7
  assert rel_err(outs[1][1], outs[0][1]) < 1e-9
  assert rel_err(outs[1][2], outs[0][2]) < 1e-5
