"""The drop-in boundary on the host side: the reference's import paths, dataset dispatch, readers."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_import_paths_resolve():
  """scripts/train.py:13-20, scripts/test.py:10-14, eval_imm.py: `imm.*` resolves to the imm_b200 modules."""
  from imm.models.imm_model import IMMModel, get_gaussian_maps           # noqa: F401
  from imm.models.base_model import BaseModel
  import imm.train.cnn_train_multi as tru
  from imm.eval import eval_imm
  from imm.utils.box import Box
  from imm.utils.dataset_import import import_dataset
  from imm.datasets.celeba_dataset import CelebADataset
  from imm.datasets.aflw_dataset import AFLWDataset
  from imm.datasets.tps_dataset import TPSDataset
  import imm_b200.models.imm_model as real
  assert IMMModel is real.IMMModel and issubclass(IMMModel, BaseModel)
  for name in ('setup_training', 'train_loop', 'train_single', 'train_multi', 'average_gradients', 'tower_loss'):
    assert callable(getattr(tru, name))
  assert callable(eval_imm.evaluate) and Box({'a': {'b': 1}}).a.b == 1
  assert import_dataset('celeba') is CelebADataset and import_dataset('aflw') is AFLWDataset
  assert issubclass(CelebADataset, TPSDataset)
  with pytest.raises(ValueError):
    import_dataset('imagenet')


def _fake_celeba(root, n=12, size=(60, 50)):
  from PIL import Image
  rng = np.random.RandomState(0)
  os.makedirs(os.path.join(root, 'Img', 'img_align_celeba_hq'))
  for d in ('Anno', 'Eval', 'MAFL'):
    os.makedirs(os.path.join(root, d))
  names = ['%06d.jpg' % (i + 1) for i in range(n)]
  pts = rng.randint(5, 45, size=(n, 10))
  for nm in names:
    arr = rng.randint(0, 255, size=(size[0], size[1], 3)).astype(np.uint8)
    Image.fromarray(arr).save(os.path.join(root, 'Img', 'img_align_celeba_hq', nm.replace('.jpg', '.png')))
    os.rename(os.path.join(root, 'Img', 'img_align_celeba_hq', nm.replace('.jpg', '.png')),
              os.path.join(root, 'Img', 'img_align_celeba_hq', nm))      # PNG bytes under the .jpg name: lossless fixture
  with open(os.path.join(root, 'Anno', 'list_landmarks_align_celeba.txt'), 'w') as f:
    f.write('%d\nheader\n' % n)
    for nm, p in zip(names, pts):
      f.write(nm + ' ' + ' '.join(str(int(v)) for v in p) + '\n')
  with open(os.path.join(root, 'Eval', 'list_eval_partition.txt'), 'w') as f:
    for i, nm in enumerate(names):
      f.write('%s %d\n' % (nm, 0 if i < 8 else 1))
  with open(os.path.join(root, 'MAFL', 'training.txt'), 'w') as f:
    f.write('\n'.join(names[8:11]) + '\n')
  with open(os.path.join(root, 'MAFL', 'testing.txt'), 'w') as f:
    f.write(names[11] + '\n')
  return names, pts.reshape(n, 5, 2).astype(np.float32)


def test_celeba_reader_lists_subsets_and_preprocesses(tmp_path):
  """celeba_dataset.py:14-92 (membership) and :136-174 (resize to R/0.8 corner-aligned, central crop, landmarks)."""
  from imm_b200.datasets.face_datasets import CelebADataset, celeba_file_list
  root = str(tmp_path)
  names, pts = _fake_celeba(root)
  _, train, kp = celeba_file_list(root, 'celeba', 'train')
  assert list(train) == names[:8] and kp.shape == (8, 5, 2)          # partition 0, minus nothing (MAFL test is image 12)
  _, val, _ = celeba_file_list(root, 'celeba', 'val')
  assert list(val) == names[8:11]                                    # partition 1 minus the MAFL test image
  _, mtrain, _ = celeba_file_list(root, 'mafl', 'train')
  assert list(mtrain) == names[8:11]                                 # round(0.1 * 3) = 0 validation images
  _, mtest, _ = celeba_file_list(root, 'mafl', 'test')
  assert list(mtest) == [names[11]]
  with pytest.raises(ValueError):
    celeba_file_list(root, 'celeba', 'test')
  ds = CelebADataset(root, 'train', dataset='celeba', tps=False, order_stream=True, image_size=[64, 64], device='cpu')
  nxt = ds.get_dataset(4, repeat=False)
  b0 = nxt()
  assert tuple(b0['image'].shape) == (4, 64, 64, 3) and tuple(b0['mask'].shape) == (4, 64, 64, 1)
  assert b0['image'].min() >= 0 and b0['image'].max() <= 255 and b0['future_image'] is b0['image']
  # landmarks: (x,y) annotation -> (y,x), scaled to the 80x80 resize, minus the 8-pixel crop margin
  want = pts[0][:, [1, 0]] * (np.array([80, 80], np.float32) / np.array([60, 50], np.float32)) - 8
  np.testing.assert_allclose(b0['landmarks'][0].numpy(), want, rtol=1e-6)
  assert nxt() is not None and nxt() is None                       # 8 images / batch 4, no repeat


def test_bench_config_table_matches_baseline():
  import json
  sys.path.insert(0, ROOT)
  import bench
  base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
  assert len(base['configs']) == 5
  for key, gb, gpus, R, K in (('c2', 64, 1, 128, 10), ('c3', 256, 8, 128, 30), ('c4', 128, 4, 128, 50), ('c5', 512, 8, 256, 10)):
    c = bench.CONFIGS[key]
    assert (c['global_batch'], c['named_gpus'], c['image_size'], c['n_maps']) == (gb, gpus, R, K)
    assert c['per_gpu'] * c['named_gpus'] == c['global_batch']
  cfg = bench.workload_config('c3', 8)
  assert cfg['global_batch'] == 256 and cfg['per_gpu_batch'] == 32 and cfg['baseline_config'] == 'c3'
