"""Alias: this module IS imm_b200.utils.tps_sampler (see imm/__init__.py)."""
import sys

import imm_b200.utils.tps_sampler as _impl

sys.modules[__name__] = _impl
