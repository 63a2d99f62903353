"""Alias: this module IS imm_b200.utils.dataset_import (see imm/__init__.py)."""
import sys

import imm_b200.utils.dataset_import as _impl

sys.modules[__name__] = _impl
