"""Alias: this module IS imm_b200.eval.eval_imm (see imm/__init__.py)."""
import sys

import imm_b200.eval.eval_imm as _impl

sys.modules[__name__] = _impl
