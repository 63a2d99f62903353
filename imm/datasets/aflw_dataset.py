"""Alias: this module IS imm_b200.datasets.face_datasets (see imm/__init__.py)."""
import sys

import imm_b200.datasets.face_datasets as _impl

sys.modules[__name__] = _impl
