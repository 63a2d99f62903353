"""Seeded synthetic inputs / VGG weights for benchmarking and smoke tests (no datasets, checkpoints or
vgg16.caffemodel.h5 are reachable offline).  Shapes and value ranges follow the reference's input contract:
image / future_image [B,R,R,3] fp32 in [0,255] un-normalised (impair_dataset.py:49); mask [B,R,R,1] =
TPSDataset._get_smooth_mask (tps_dataset.py:47-67; margin 10, step 20, celeba_dataset.py:165)."""
import math

import numpy as np
import torch
import torch.nn.functional as F


def smooth_mask(h, w, margin=10, step=20, b=0.4):
  def smooth_step(n, bb):
    x = torch.linspace(-1.0, 1.0, n, dtype=torch.float32)
    return 0.5 + 0.5 * torch.tanh(x / bb)

  def strip(size):
    return torch.cat([torch.zeros(margin), smooth_step(step, b), torch.ones(size - 2 * margin - 2 * step),
                      smooth_step(step, -b), torch.zeros(margin)])
  return strip(h)[:, None] * strip(w)[None]


def synthetic_inputs(batch, image_size=128, seed=0, pin=False):
  """CPU tensors: smooth random image pairs (uniform noise at 1/8 resolution, bilinearly upsampled) + mask."""
  gen = torch.Generator().manual_seed(1000 + seed)

  def smooth_image():
    lo = torch.rand((batch, 3, image_size // 8, image_size // 8), generator=gen) * 255.0
    hi = F.interpolate(lo, size=(image_size, image_size), mode='bilinear', align_corners=False)
    return hi.permute(0, 2, 3, 1).contiguous()
  image, future = smooth_image(), smooth_image()
  mask = smooth_mask(image_size, image_size).view(1, image_size, image_size, 1).repeat(batch, 1, 1, 1).contiguous()
  out = {'image': image, 'future_image': future, 'mask': mask}
  if pin:
    out = {k: v.pin_memory() for k, v in out.items()}
  return out


VGG_CONVS = [('conv1_1', 64), ('conv1_2', 64), ('conv2_1', 128), ('conv2_2', 128), ('conv3_1', 256),
             ('conv3_2', 256), ('conv3_3', 256), ('conv4_1', 512), ('conv4_2', 512), ('conv4_3', 512),
             ('conv5_1', 512), ('conv5_2', 512), ('conv5_3', 512)]


def synthetic_vgg_caffe_dict(seed=1):
  """Seeded He-normal weights in the dict layout deepdish returns for vgg16.caffemodel.h5
  (build_vgg16.py:16; vgg16.py:19-40,76-87): data[name]['0'] = W [O,I,3,3], ['1'] = bias [O];
  data['batch_'+name]['0'|'1'|'2'] = mean*s, var*s, s (Caffe BatchNorm blobs, grayscale input: Cin=1)."""
  rng = np.random.RandomState(seed)
  data, cin = {}, 1
  for name, cout in VGG_CONVS:
    std = math.sqrt(2.0 / (9 * cin))
    data[name] = {'0': (rng.randn(cout, cin, 3, 3) * std).astype(np.float32),
                  '1': (rng.randn(cout) * 0.05).astype(np.float32)}
    s = np.float32(2.0)
    data['batch_' + name] = {'0': (rng.randn(cout) * 0.1).astype(np.float32) * s,
                             '1': (0.5 + rng.rand(cout)).astype(np.float32) * s,
                             '2': np.array([s], dtype=np.float32)}
    cin = cout
  return data
