"""Alias: this module IS imm_b200.models.imm_model (see imm/__init__.py)."""
import sys

import imm_b200.models.imm_model as _impl

sys.modules[__name__] = _impl
