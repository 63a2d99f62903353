"""Alias: this module IS imm_b200.datasets.tps_dataset (see imm/__init__.py)."""
import sys

import imm_b200.datasets.tps_dataset as _impl

sys.modules[__name__] = _impl
