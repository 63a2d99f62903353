"""Loader for the colorization-pretrained VGG16 (`vgg16.caffemodel.h5`, configs/paths/default.yaml:6).
The reference reads it with deepdish (build_vgg16.py:16: dd.io.load(file, '/data')); deepdish/h5py are not
installed in this image and the file is a download, so this only works where h5py is available."""


def load_caffe_h5(path):
  try:
    import h5py
  except ImportError:
    raise RuntimeError('h5py is not installed: cannot read %s; pass a Caffe-layout dict to IMMModel.load_vgg() '
                       '(imm_b200.utils.synthetic.synthetic_vgg_caffe_dict for benchmarking)' % path)
  data = {}
  with h5py.File(path, 'r') as f:
    root = f['data']
    for name in root:
      data[name] = {k: root[name][k][()] for k in root[name]}
  return data
