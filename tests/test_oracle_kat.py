"""Hand-derived known-answer tests that pin every TF-1.10 semantic ([TF-sem]) the
oracle restates.  Expected values are worked out by hand in the comments, not produced by the oracle."""
import math

import numpy as np
import torch

from oracle import imm_oracle as O


def test_same_pad_rule():
  # stride 1, k=3: 1/1; k=7: 3/3.  stride 2, k=3, even input: total 1 -> 0 before, 1 after.
  assert O.same_pad(128, 3, 1) == (1, 1)
  assert O.same_pad(128, 7, 1) == (3, 3)
  assert O.same_pad(128, 3, 2) == (0, 1)
  assert O.same_pad(16, 1, 1) == (0, 0)
  # odd input, stride 2, k=3: out=ceil(5/2)=3, total=(3-1)*2+3-5=2 -> 1/1
  assert O.same_pad(5, 3, 2) == (1, 1)


def test_conv_stride2_same_is_asymmetric():
  # 4x4 input = row index*10 + col index; all-ones 3x3 kernel, stride 2 -> 2x2 output.
  # window of out (i,j) covers rows 2i..2i+2, cols 2j..2j+2 (zero beyond index 3).
  x = torch.tensor([[r * 10.0 + c for c in range(4)] for r in range(4)]).view(1, 4, 4, 1)
  w = torch.ones(3, 3, 1, 1)
  y = O.conv2d_same(x, w, None, 2)[0, :, :, 0]
  # out(0,0) = rows 0-2, cols 0-2: sum = 3*(0+10+20) + 3*(0+1+2) = 90 + 9 = 99
  # out(0,1) = rows 0-2, cols 2-3 (col 4 is pad): (2+3)+(12+13)+(22+23) = 75
  # out(1,0) = rows 2-3, cols 0-2: (20+21+22)+(30+31+32) = 156
  # out(1,1) = rows 2-3, cols 2-3: 22+23+32+33 = 110
  assert y.tolist() == [[99.0, 75.0], [156.0, 110.0]]


def test_conv_is_cross_correlation_hwio():
  # delta input at (1,1); kernel value = 10*r + s -> output(h,w) = w[1-h+1, 1-w+1] flipped pattern of
  # cross-correlation: y[h,w] = sum_rs x[h+r-1, w+s-1] w[r,s]  => y[h,w] = w[2-h, 2-w] for the delta at (1,1).
  x = torch.zeros(1, 3, 3, 1)
  x[0, 1, 1, 0] = 1.0
  w = torch.tensor([[10.0 * r + s for s in range(3)] for r in range(3)]).view(3, 3, 1, 1)
  y = O.conv2d_same(x, w, None, 1)[0, :, :, 0]
  expect = [[w[2 - h, 2 - ww, 0, 0].item() for ww in range(3)] for h in range(3)]
  assert y.tolist() == expect
  # HWIO channel semantics: 2 in, 2 out, 1x1
  x2 = torch.tensor([1.0, 2.0]).view(1, 1, 1, 2)
  w2 = torch.tensor([[3.0, 4.0], [5.0, 6.0]]).view(1, 1, 2, 2)   # [ci, co]
  y2 = O.conv2d_same(x2, w2, torch.tensor([0.5, -0.5]), 1).view(-1)
  assert y2.tolist() == [1 * 3 + 2 * 5 + 0.5, 1 * 4 + 2 * 6 - 0.5]


def test_legacy_bilinear_x2():
  # [0,1,2,3] -> [0,.5,1,1.5,2,2.5,3,3]  (src = dst*0.5; last sample clamps)
  x = torch.tensor([0.0, 1.0, 2.0, 3.0]).view(1, 1, 4, 1)
  y = O.resize_bilinear(x, [1, 8]).view(-1)
  assert y.tolist() == [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.0]
  yh = O.resize_bilinear(x.permute(0, 2, 1, 3), [8, 1]).view(-1)
  assert yh.tolist() == y.tolist()


def test_legacy_bilinear_integer_downscale_is_subsample():
  x = torch.arange(64, dtype=torch.float32).view(1, 8, 8, 1)
  y = O.resize_bilinear(x, [2, 2])[0, :, :, 0]
  assert y.tolist() == [[0.0, 4.0], [32.0, 36.0]]       # x[::4, ::4]


def test_align_corners_resize():
  # 4 -> 2 with align_corners: scale=(4-1)/(2-1)=3 -> samples at 0 and 3
  x = torch.tensor([0.0, 1.0, 2.0, 3.0]).view(1, 1, 4, 1)
  assert O.resize_bilinear(x, [1, 2], align_corners=True).view(-1).tolist() == [0.0, 3.0]
  # 3 -> 5: scale = 2/4 = 0.5 -> [0, .5, 1, 1.5, 2]
  x = torch.tensor([0.0, 1.0, 2.0]).view(1, 1, 3, 1)
  assert O.resize_bilinear(x, [1, 5], align_corners=True).view(-1).tolist() == [0.0, 0.5, 1.0, 1.5, 2.0]


def test_fused_bn_training_and_moving_stats():
  # channel values [1,2,3,4] (N=4): mean 2.5, biased var 1.25, Bessel var 1.25*4/3 = 5/3
  x = torch.tensor([1.0, 2.0, 3.0, 4.0]).view(1, 2, 2, 1)
  g, b = torch.tensor([2.0]), torch.tensor([0.5])
  y, mm, mv = O.batch_norm(x, g, b, torch.zeros(1), torch.ones(1), training=True)
  inv = 1.0 / math.sqrt(1.25 + 1e-3)
  expect = [(v - 2.5) * inv * 2.0 + 0.5 for v in [1.0, 2.0, 3.0, 4.0]]
  np.testing.assert_allclose(y.view(-1).numpy(), expect, rtol=1e-6)
  np.testing.assert_allclose(mm.item(), 0.0 - (0.0 - 2.5) * 0.01, rtol=1e-6)           # 0.025
  np.testing.assert_allclose(mv.item(), 1.0 - (1.0 - 5.0 / 3.0) * 0.01, rtol=1e-6)     # 1.006667
  # inference uses the moving stats
  y2, _, _ = O.batch_norm(x, g, b, torch.tensor([1.0]), torch.tensor([4.0]), training=False)
  np.testing.assert_allclose(y2.view(-1).numpy(), [(v - 1.0) / math.sqrt(4.001) * 2 + 0.5 for v in [1, 2, 3, 4]], rtol=1e-6)


def test_tf_adam_scalar_step():
  # g=0.5, lr=1e-3, t=1: m=.05, v=2.5e-4; lr_t = 1e-3*sqrt(.001)/.1; step = lr_t*m/(sqrt(v)+1e-8)
  var, m, v = O.adam_step(torch.tensor(1.0), torch.tensor(0.5), torch.tensor(0.0), torch.tensor(0.0), 1e-3, 1)
  lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
  expect = 1.0 - lr_t * 0.05 / (math.sqrt(2.5e-4) + 1e-8)
  np.testing.assert_allclose(var.item(), expect, rtol=1e-6)
  np.testing.assert_allclose(m.item(), 0.05, rtol=1e-6)
  np.testing.assert_allclose(v.item(), 2.5e-4, rtol=1e-6)
  # tiny gradient: eps matters and is OUTSIDE the bias correction (differs from torch.optim.Adam)
  var2, _, _ = O.adam_step(torch.tensor(0.0), torch.tensor(1e-9), torch.tensor(0.0), torch.tensor(0.0), 1e-3, 1)
  expect2 = -lr_t * 1e-10 / (math.sqrt(1e-21) + 1e-8)
  np.testing.assert_allclose(var2.item(), expect2, rtol=1e-5)


def test_clip_by_norm():
  g = torch.tensor([1.2, 1.6])            # norm 2 -> scaled to norm 1
  np.testing.assert_allclose(O.clip_by_norm(g, 1.0).numpy(), [0.6, 0.8], rtol=1e-6)
  g = torch.tensor([0.3, 0.4])            # norm .5 < 1 -> unchanged
  np.testing.assert_allclose(O.clip_by_norm(g, 1.0).numpy(), [0.3, 0.4], rtol=1e-6)
  assert O.clip_by_norm(torch.zeros(3), 1.0).tolist() == [0.0, 0.0, 0.0]


def test_learning_rate_schedule_and_global_step_quirk():
  assert O.learning_rate(0) == 1e-3
  assert O.learning_rate(99999) == 1e-3
  np.testing.assert_allclose(O.learning_rate(100000), 0.95e-3, rtol=1e-12)
  np.testing.assert_allclose(O.learning_rate(-1), 1e-3 / 0.95, rtol=1e-12)   # train.py default global_step=-1


def test_get_coord_is_marginal_softmax_not_2d():
  # 2x2 map, K=1: x = [[0, 2],[4, 6]].  rows: mean over w -> [1, 5] -> softmax -> p=[s(-4), s(4)],
  # mu_y = -p0 + p1 = tanh(2).   cols: mean over h -> [2, 4] -> mu_x = tanh(1).
  x = torch.tensor([[0.0, 2.0], [4.0, 6.0]]).view(1, 2, 2, 1)
  gy, py = O.get_coord(x, 2, 2)
  gx, px = O.get_coord(x, 1, 2)
  np.testing.assert_allclose(gy.item(), math.tanh(2.0), rtol=1e-6)
  np.testing.assert_allclose(gx.item(), math.tanh(1.0), rtol=1e-6)


def test_gaussian_maps_rot():
  mu = torch.tensor([[[0.0, -1.0]]])      # (y, x): centre row, left column
  g = O.get_gaussian_maps(mu, [3, 3], 10.0, mode='rot')[0, :, :, 0]   # grid {-1,0,1}
  # G[i,j] = exp(-((y_i-0)^2 + (x_j+1)^2)*100)
  expect = [[math.exp(-(yy ** 2 + (xx + 1) ** 2) * 100) for xx in (-1, 0, 1)] for yy in (-1, 0, 1)]
  np.testing.assert_allclose(g.numpy(), expect, rtol=1e-5, atol=1e-30)
  assert g[1, 0].item() == 1.0


def test_max_pool():
  x = torch.tensor([[1.0, 5.0, 2.0, 0.0], [3.0, 4.0, 9.0, 1.0], [0.0, 0.0, 7.0, 7.0], [0.0, 0.0, 7.0, 8.0]]).view(1, 4, 4, 1)
  assert O.max_pool_2x2(x)[0, :, :, 0].tolist() == [[5.0, 9.0], [0.0, 8.0]]


def test_vgg_bn_fold():
  data = {'conv1_1': {'0': np.ones((2, 1, 3, 3), np.float32), '1': np.array([1.0, 2.0], np.float32)},
          'batch_conv1_1': {'0': np.array([2.0, 4.0], np.float32), '1': np.array([8.0, 2.0], np.float32),
                            '2': np.array([2.0], np.float32)}}
  full = O.synthetic_vgg_caffe_dict(3)
  full.update(data)
  p = O.load_vgg_params(full)
  W = p['SelfSupReconstructionLoss/vgg16/conv1_1/weights']
  b = p['SelfSupReconstructionLoss/vgg16/conv1_1/biases']
  assert tuple(W.shape) == (3, 3, 1, 2)
  # sigma = sqrt(1e-5 + var*s/s) = sqrt(1e-5+4), sqrt(1e-5+1); mu = 1, 2
  s0, s1 = math.sqrt(4 + 1e-5), math.sqrt(1 + 1e-5)
  np.testing.assert_allclose(W[0, 0, 0].numpy(), [1 / s0, 1 / s1], rtol=1e-6)
  np.testing.assert_allclose(b.numpy(), [(1 - 1) / s0, (2 - 2) / s1], atol=1e-7)


def test_perceptual_level_formula_and_agg_update():
  # one level ('input'), no mask: s = mean(d^2); wl = a + .01(s-a); L = s/wl; loss = 1000*L
  st = O.State(n_maps=2, perceptual_comp=('input',))
  st.buffers['SelfSupReconstructionLoss/input_agg'] = torch.tensor(100.0)
  st.vgg = O.load_vgg_params(O.synthetic_vgg_caffe_dict(1))
  gt = torch.full((1, 16, 16, 3), 10.0)
  pr = torch.full((1, 16, 16, 3), 7.0)
  loss, agg, _ = O.perceptual_loss(st, gt, pr, None, True)
  s = 9.0
  wl = 100.0 + 0.01 * (s - 100.0)
  np.testing.assert_allclose(loss.item(), 1000.0 * s / wl, rtol=1e-6)
  np.testing.assert_allclose(agg['SelfSupReconstructionLoss/input_agg'].item(), wl, rtol=1e-6)


def test_smooth_mask_profile():
  m = O.smooth_mask(128, 128)
  assert m.shape == (128, 128)
  assert float(m[:10].abs().max()) == 0.0 and float(m[-10:].abs().max()) == 0.0
  assert float(m[64, 64]) == 1.0
  # first ramp sample: 0.5 + 0.5*tanh(-1/0.4)
  np.testing.assert_allclose(float(m[10, 64]), 0.5 + 0.5 * math.tanh(-2.5), rtol=1e-5)


def test_param_inventory_matches_survey():
  st = O.init_state(O.State(n_maps=10), seed=0)
  assert len(st.params) == 96
  assert sum(p.numel() for p in st.params.values()) == 4138067            # "4.14 M" 
  assert tuple(st.params['model/renderer/conv_1/conv_1/w'].shape) == (3, 3, 266, 256)
  assert tuple(st.params['model/pose_encoder/conv_1/conv_1/w'].shape) == (1, 1, 256, 10)
  assert 'model/renderer/conv_8/batch_normalization/gamma' not in st.params
  assert tuple(st.params['model/renderer/conv_8/conv_8/w'].shape) == (3, 3, 32, 9)
  assert [r[0] for r in O.renderer_spec(32, 256, 9)][-1] == 'conv_10'
