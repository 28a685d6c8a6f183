"""Profiling driver (GPU): scripts/run_stage.py with the one-launch cluster form of the small-N block (pf_stage.cu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from polyphonicformer_b200 import _cabi  # noqa: E402

_cabi.load().pf_set_fused_update(1)
exec(open(os.path.join(ROOT, 'scripts', 'run_stage.py')).read())
