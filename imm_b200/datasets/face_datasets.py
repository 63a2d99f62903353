"""CelebA / MAFL and AFLW readers -- host side of SURVEY 8(f) row N1, mirroring imm/datasets/celeba_dataset.py and
aflw_dataset.py: the same directory contract, subsets, preprocessing and output dict; decoding is PIL on the host, the
two thin-plate-spline warps run on the GPU (TPSDataset / immb_tps_warp).

Directory contract (celeba_dataset.py:14-92):  <root>/Img/img_align_celeba_hq/*.jpg,
<root>/Anno/list_landmarks_align_celeba.txt (two header lines, then `name x1 y1 .. x5 y5`),
<root>/Eval/list_eval_partition.txt (`name {0,1,2}`), <root>/MAFL/{training,testing}.txt.
AFLW (aflw_dataset.py:14-39): <root>/aflw_{train,test}_images.txt, aflw_{train,test}_keypoints.mat ('gt' [N,5,2] as
(x,y) -> stored (y,x), 'hw'), images under <root>/output/.
Preprocessing: CelebA resizes to round(R/0.8) with a corner-aligned bilinear resize and takes the central RxR crop
(celeba_dataset.py:136-174); AFLW resizes straight to RxR (aflw_dataset.py:80-112).  Output per batch:
image / future_image [B,R,R,3] fp32 in [0,255], mask [B,R,R,1], landmarks / future_landmarks [B,5,2] (y,x) pixels."""
import os

import numpy as np
import torch
import torch.nn.functional as F

from .tps_dataset import TPSDataset
from ..utils.tps_sampler import apply_tps


def _read_lines(path):
  with open(path, 'r') as f:
    return f.read().splitlines()


def celeba_file_list(data_root, dataset, subset):
  """-> (image_dir, file names [n], keypoints [n,5,2] as (x,y) pixel pairs in annotation order)."""
  rows = [ln.split() for ln in _read_lines(os.path.join(data_root, 'Anno', 'list_landmarks_align_celeba.txt'))[2:]]
  names = np.array([r[0] for r in rows])
  pts = np.array([[int(v) for v in r[1:]] for r in rows], dtype=np.float32).reshape(-1, 5, 2)
  mafl_train = set(_read_lines(os.path.join(data_root, 'MAFL', 'training.txt')))
  mafl_test = set(_read_lines(os.path.join(data_root, 'MAFL', 'testing.txt')))
  in_train = np.array([n in mafl_train for n in names])
  in_test = np.array([n in mafl_test for n in names])
  # membership codes: 1 train, 2 val (celeba partition 1), 3 celeba test, 4 MAFL test, 5 = last 10 % of MAFL train
  code = np.zeros(len(names), dtype=np.int32)
  if dataset == 'celeba':
    part = [int(ln.split()[1]) for ln in _read_lines(os.path.join(data_root, 'Eval', 'list_eval_partition.txt'))]
    code[:] = np.asarray(part, dtype=np.int32) + 1
    wanted = {'train': 1, 'val': 2}
  elif dataset == 'mafl':
    code[in_train] = 1
    wanted = {'train': 1, 'test': 4, 'train10': 5}
  else:
    raise ValueError('Dataset = %s not recognized.' % dataset)
  code[in_test] = 4
  train_idx = np.nonzero(in_train)[0]
  n_val = int(round(0.1 * len(train_idx)))
  if n_val:
    code[train_idx[-n_val:]] = 5
  if subset not in wanted:
    raise ValueError('subset = %s for %s dataset not recognized.' % (subset, dataset))
  keep = code == wanted[subset]
  return os.path.join(data_root, 'Img', 'img_align_celeba_hq'), names[keep], pts[keep]


def aflw_file_list(data_dir, subset):
  """-> (image_dir, file names, keypoints [n,5,2] (y,x), sizes [n,2] (h,w))."""
  from scipy.io import loadmat
  part = 'train' if subset in ('train', 'val') else 'test'
  names = _read_lines(os.path.join(data_dir, 'aflw_%s_images.txt' % part))
  mat = loadmat(os.path.join(data_dir, 'aflw_%s_keypoints.mat' % part))
  kp, hw = mat['gt'][:, :, [1, 0]], mat['hw']
  if part == 'train':
    n_val = int(round(0.1 * len(names)))
    sl = slice(0, len(names) - n_val) if subset == 'train' else slice(len(names) - n_val, len(names))
    names, kp, hw = names[sl], kp[sl], hw[sl]
  return os.path.join(data_dir, 'output'), np.array(names), kp, hw


class _FaceDataset(TPSDataset):
  """Shared reader: file list -> decoded, resized batches (+ landmarks) -> optional GPU TPS pair generation."""
  LANDMARK_LABELS = {'left_eye': 0, 'right_eye': 1}
  N_LANDMARKS = 5

  def __init__(self, data_dir, subset, max_samples=None, image_size=[128, 128], order_stream=False, landmarks=False,
               tps=True, device='cuda:0', seed=0, **tps_kwargs):
    super(_FaceDataset, self).__init__(data_dir, subset, max_samples=max_samples, image_size=image_size,
                                       order_stream=order_stream, landmarks=landmarks, tps=tps, device=device,
                                       seed=seed, **tps_kwargs)
    self._data_dir, self._subset, self._max_samples, self._order_stream = data_dir, subset, max_samples, order_stream
    self.image_size = image_size

  # subclasses: self._image_dir, self._images, self._keypoints_yx(idx, original_hw) and self._resize(image)
  def __len__(self):
    n = len(self._images)
    return n if self._max_samples is None else min(n, int(self._max_samples))

  def _decode(self, idx):
    from PIL import Image
    with Image.open(os.path.join(self._image_dir, str(self._images[idx]))) as im:
      arr = np.asarray(im.convert('RGB'), dtype=np.float32)
    return torch.from_numpy(arr)                                  # [h,w,3] in [0,255]

  @staticmethod
  def _resize_ac(image_hwc, out_hw):
    """tf.image.resize_images(..., BILINEAR, align_corners=True)."""
    x = image_hwc.permute(2, 0, 1).unsqueeze(0)
    return F.interpolate(x, size=tuple(int(v) for v in out_hw), mode='bilinear', align_corners=True)[0].permute(1, 2, 0)

  @staticmethod
  def _resize_points(points_yx, size_hw, new_size_hw):
    """impair_dataset.py `_resize_points`: scale by new/old per axis."""
    return points_yx * (np.asarray(new_size_hw, np.float32) / np.asarray(size_hw, np.float32))

  def sample(self, idx):
    raise NotImplementedError

  def get_dataset(self, batch_size, repeat=True, shuffle=False, num_preprocess_threads=12, rank=0, world=1):
    R = self._image_size[0]
    mask = self._get_smooth_mask(R, R, 10, 20).view(1, R, R, 1).repeat(batch_size, 1, 1, 1)      # celeba_dataset.py:165
    n = len(self)
    rng = np.random.RandomState(self._seed + rank)
    state = {'pos': rank * batch_size, 'order': np.arange(n) if self._order_stream else rng.permutation(n)}

    def next_batch():
      if state['pos'] + batch_size > n:
        if not repeat:
          return None
        state['pos'] = rank * batch_size
        if not self._order_stream:
          state['order'] = rng.permutation(n)
      idx = state['order'][state['pos']:state['pos'] + batch_size]
      state['pos'] += batch_size * world
      ims, lms = zip(*[self.sample(int(i)) for i in idx])
      img = torch.stack(ims).contiguous()
      lm = torch.from_numpy(np.stack(lms).astype(np.float32))
      out = {'image': img, 'future_image': img, 'mask': mask, 'landmarks': lm, 'future_landmarks': lm}
      if self._tps:
        dev = self._device
        pair = apply_tps(img.to(dev, non_blocking=True), mask.to(dev), self._target_sampler, self._source_sampler)
        out.update(pair)          # landmarks are not warped (the reference refuses landmarks together with TPS)
      return out
    return next_batch


class CelebADataset(_FaceDataset):
  """imm/datasets/celeba_dataset.py:95-174."""

  def __init__(self, data_dir, subset, dataset=None, name='CelebADataset', **kwargs):
    super(CelebADataset, self).__init__(data_dir, subset, **kwargs)
    assert dataset is not None
    self._dataset = dataset
    self._image_dir, self._images, self._keypoints = celeba_file_list(data_dir, dataset, subset)

  def sample(self, idx):
    R = self._image_size[0]
    resize_sz = int(np.round(R / 0.8))
    margin = int(np.round((resize_sz - R) / 2.0))
    im = self._decode(idx)
    lm_yx = self._keypoints[idx][:, [1, 0]]                     # annotation is (x,y); the model works in (y,x)
    lm_yx = self._resize_points(lm_yx, im.shape[:2], [resize_sz, resize_sz]) - margin
    im = self._resize_ac(im, [resize_sz, resize_sz])[margin:margin + R, margin:margin + R]
    return im.contiguous(), lm_yx


class AFLWDataset(_FaceDataset):
  """imm/datasets/aflw_dataset.py:42-123."""

  def __init__(self, data_dir, subset, name='AFLWDataset', **kwargs):
    super(AFLWDataset, self).__init__(data_dir, subset, **kwargs)
    self._image_dir, self._images, self._keypoints, self._sizes = aflw_file_list(data_dir, subset)

  def sample(self, idx):
    R = self._image_size[0]
    im = self._decode(idx)
    lm_yx = self._keypoints[idx][:, [1, 0]]                     # aflw_dataset.py:115-117 (undoes the swap at load)
    lm_yx = self._resize_points(lm_yx, self._sizes[idx], [R, R])
    return self._resize_ac(im, [R, R]).contiguous(), lm_yx
