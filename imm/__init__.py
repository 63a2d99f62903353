"""`imm` -- the reference's package name, served by imm_b200.

The reference's callers (scripts/train.py:13-20, scripts/test.py:10-14, examples/visualize.ipynb) import
`imm.models.imm_model.IMMModel`, `imm.train.cnn_train_multi`, `imm.eval.eval_imm`, `imm.utils.box.Box`,
`imm.utils.dataset_import.import_dataset`, `imm.datasets.*`.  Every module below is the imm_b200 module of the same
role (the very same module object, so classes compare identical), which makes this repository a drop-in for those
imports.  Modules of the reference that are TensorFlow graph helpers with no role on the CUDA path (imm.tf_utils,
imm.models.selfsup, imm.data_utils) are intentionally absent: INTEGRATION.md lists what replaces each."""
