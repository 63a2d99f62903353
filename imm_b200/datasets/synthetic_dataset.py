"""Synthetic image-pair stream with the reference's input contract: dict(image, future_image
[B,R,R,3] fp32 in [0,255], mask [B,R,R,1]).  Stands in for CelebADataset/AFLWDataset + TPSDataset (datasets and the
TPS warp are out of scope for this path).  Batches are pre-generated in pinned host memory and cycled."""
from ..utils.synthetic import synthetic_inputs


class SyntheticDataset(object):
  def __init__(self, data_dir=None, subset='train', image_size=(128, 128), n_batches=4, seed=0, **unused):
    self.image_size = image_size[0] if isinstance(image_size, (tuple, list)) else image_size
    self.n_batches, self.seed = n_batches, seed

  def get_dataset(self, batch_size, repeat=True, shuffle=False, num_preprocess_threads=12, rank=0):
    batches = [synthetic_inputs(batch_size, self.image_size, seed=self.seed + 100 * rank + i, pin=True)
               for i in range(self.n_batches)]
    state = {'i': 0}

    def next_batch():
      b = batches[state['i'] % len(batches)]
      state['i'] += 1
      return b
    return next_batch
