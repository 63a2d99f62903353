"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/imm_oracle.py header; PARITY UNPINNED: the reference has no tests).

CPU restatement of the reference's random thin-plate-spline warp:
imm/utils/tps_sampler.py (TPSGridGen :111-165, sample_tps_w :168-189, TPSRandomSampler.forward :77-98) and
imm/datasets/tps_dataset.py:_apply_tps (:70-96).  The reference runs PyTorch 0.4.1, whose F.grid_sample is
bilinear / zero padding with what later became `align_corners=True`; that is made explicit here."""
import numpy as np
import torch
import torch.nn.functional as F


def tps_L(Ho, Wo, Hc, Wc):
  """TPSGridGen.__init__ (tps_sampler.py:111-148): L [Ho*Wo, Hc*Wc+3] = [U(|g - c|^2), 1, g_x, g_y], U(d2) = d2 log d2."""
  xx, yy = np.meshgrid(np.linspace(-1, 1, Wo), np.linspace(-1, 1, Ho))
  grid = np.c_[xx.flatten(), yy.flatten()].astype(np.float32)
  xx, yy = np.meshgrid(np.linspace(-1, 1, Wc), np.linspace(-1, 1, Hc))
  cp = np.c_[xx.flatten(), yy.flatten()].astype(np.float32)
  d = grid[:, None, :].astype(np.float64) - cp[None, :, :].astype(np.float64)
  Dx = (d ** 2).sum(-1)                               # scipy cdist 'sqeuclidean'
  Dx = np.clip(Dx, 1e-8, None)
  Kp = np.log(Dx) * Dx
  L = np.c_[Kp, np.ones((grid.shape[0], 1), dtype=np.float32), grid]
  return torch.from_numpy(L.astype(np.float32))


def tps_grid(w_tps, Ho, Wo, Hc, Wc):
  """TPSGridGen.forward (:151-165): [B, M+3, 2] -> [B, Ho, Wo, 2] (x, y) sampling coordinates."""
  L = tps_L(Ho, Wo, Hc, Wc)
  g = torch.matmul(L, w_tps.float())
  return g.reshape(w_tps.shape[0], Ho, Wo, 2)


def sample_tps_w(Hc, Wc, warpsd, rotsd, scalesd, transsd, rng=np.random):
  """tps_sampler.py:168-189, same draw order; `rng` = a numpy RandomState (the reference uses the global one)."""
  Nc = Hc * Wc
  mask = (rng.rand(Nc, 2) > 0.5).astype(np.float32)
  W = warpsd[0] * rng.randn(Nc, 2) + warpsd[1] * (mask * rng.randn(Nc, 2))
  rnd = rng.randn
  rot = np.deg2rad(rnd() * rotsd)
  sc = 1.0 + rnd() * scalesd
  aff = [[transsd * rnd(), transsd * rnd()], [sc * np.cos(rot), sc * -np.sin(rot)], [sc * np.sin(rot), sc * np.cos(rot)]]
  return np.r_[W, aff]


def tps_warp_nhwc(x, w_tps, Hc=10, Wc=10):
  """TPSRandomSampler.forward with pad=False (tps_dataset.py:37-45 constructs both samplers with pad=False):
  grid_sample(input, grid) -- bilinear, zeros padding, corner-aligned.  x [B,H,W,C] -> [B,H,W,C]."""
  B, H, W, C = x.shape
  grid = tps_grid(w_tps, H, W, Hc, Wc)
  out = F.grid_sample(x.permute(0, 3, 1, 2).float(), grid, mode='bilinear', padding_mode='zeros', align_corners=True)
  return out.permute(0, 2, 3, 1)


def apply_tps(image, mask, w_target, w_source):
  """TPSDataset._apply_tps (tps_dataset.py:70-96): [mask ; image] is warped by the target sampler -> future pair;
  the WARPED tensor is warped again by the source sampler -> source pair.  Returns the `inputs` dict of SURVEY 8a row 0."""
  x = torch.cat([mask, image], dim=3)
  fut = tps_warp_nhwc(x, w_target)
  src = tps_warp_nhwc(fut, w_source)
  return {'image': src[..., 1:], 'future_image': fut[..., 1:], 'mask': fut[..., 0:1], 'source_mask': src[..., 0:1]}
