"""import_dataset(name) -> dataset class, the role of imm/utils/dataset_import.py (scripts/train.py:116,
scripts/test.py:92-118).  Names the reference's configs use (`training.dset`): 'celeba', 'aflw'; plus 'synthetic' (seeded
pair stream, no files) and 'tps' (synthetic base images through the GPU TPS warps).  Unknown names raise."""


def import_dataset(dataset_name):
  name = str(dataset_name).lower()
  if name == 'celeba':
    from ..datasets.face_datasets import CelebADataset
    return CelebADataset
  if name == 'aflw':
    from ..datasets.face_datasets import AFLWDataset
    return AFLWDataset
  if name == 'synthetic':
    from ..datasets.synthetic_dataset import SyntheticDataset
    return SyntheticDataset
  if name == 'tps':
    from ..datasets.tps_dataset import TPSDataset
    return TPSDataset
  raise ValueError('Dataset %r is not known (celeba | aflw | synthetic | tps)' % (dataset_name,))
