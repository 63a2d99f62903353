"""Attribute-access config dictionary + YAML loader with ${name} interpolation.

Stands in for the reference's vendored `Box` (imm/utils/box.py) and `metayaml.read`
(scripts/train.py:43-51): several YAML files are merged in order, `${key}` / `${a.b}` references are
resolved against the merged document, and the result supports both `cfg.model.n_maps` and
`cfg['model']['n_maps']` plus `hasattr(cfg, 'key')` probing (imm_model.py:285,349)."""
import re

import yaml


class Box(dict):
  def __init__(self, *args, **kwargs):
    super(Box, self).__init__()
    for k, v in dict(*args, **kwargs).items():
      self[k] = v

  @staticmethod
  def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, Box):
      return Box(v)
    if isinstance(v, list):
      return [Box._wrap(x) for x in v]
    return v

  def __setitem__(self, k, v):
    super(Box, self).__setitem__(k, Box._wrap(v))

  def __getattr__(self, k):
    try:
      return self[k]
    except KeyError:
      raise AttributeError(k)

  def __setattr__(self, k, v):
    self[k] = v

  def to_dict(self):
    return {k: (v.to_dict() if isinstance(v, Box) else v) for k, v in self.items()}


def _merge(dst, src):
  for k, v in src.items():
    if isinstance(v, dict) and isinstance(dst.get(k), dict):
      _merge(dst[k], v)
    else:
      dst[k] = v
  return dst


_REF = re.compile(r'\$\{([A-Za-z0-9_.]+)\}')


def _lookup(root, path):
  cur = root
  for part in path.split('.'):
    cur = cur[part]
  return cur


def _resolve(node, root, depth=0):
  if depth > 16:
    raise ValueError('config interpolation too deep (cycle?)')
  if isinstance(node, dict):
    return {k: _resolve(v, root, depth) for k, v in node.items()}
  if isinstance(node, list):
    return [_resolve(v, root, depth) for v in node]
  if isinstance(node, str):
    m = _REF.fullmatch(node)
    if m:
      return _resolve(_lookup(root, m.group(1)), root, depth + 1)
    if _REF.search(node):
      return _resolve(_REF.sub(lambda mm: str(_resolve(_lookup(root, mm.group(1)), root, depth + 1)), node), root,
                      depth + 1)
  return node


def read_configs(file_names):
  """metayaml.read(file_names) equivalent for the reference's configs (configs/**/*.yaml)."""
  if isinstance(file_names, str):
    file_names = [file_names]
  merged = {}
  for fn in file_names:
    with open(fn, 'r') as f:
      doc = yaml.safe_load(f) or {}
    _merge(merged, doc)
  return Box(_resolve(merged, merged))


def default_model_config(n_maps=10):
  """The `model` section shared by all six shipped experiment configs (celeba-10pts.yaml:25-46)."""
  return Box({'gauss_std': 0.10, 'gauss_mode': 'rot', 'n_maps': n_maps, 'n_filters': 32,
              'block_sizes': [1, 1, 1], 'n_filters_render': 32, 'renderer_stride': 2, 'min_res': 16,
              'same_n_filt': False, 'reconstruction_loss': 'perceptual',
              'perceptual': {'l2': True, 'comp': ['input', 'conv1_2', 'conv2_2', 'conv3_2', 'conv4_2', 'conv5_2'],
                             'net_file': 'data/models/vgg16.caffemodel.h5'},
              'loss_mask': True, 'confidence': False, 'channels_bug_fix': True})
