"""Alias: this module IS imm_b200.train.cnn_train_multi (see imm/__init__.py)."""
import sys

import imm_b200.train.cnn_train_multi as _impl

sys.modules[__name__] = _impl
