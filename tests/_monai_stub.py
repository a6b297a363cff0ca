"""The MONAI stand-in lives in baseline/monai_stub.py (shared with bench.py's model-level legs); re-exported here for
the tests and tests/golden/make_golden.py."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.path.join(_ROOT, "baseline") not in sys.path:
    sys.path.insert(0, os.path.join(_ROOT, "baseline"))
from monai_stub import install, reference_root  # noqa: E402,F401
