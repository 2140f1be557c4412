"""Seeded synthetic inputs live in the repository-root module ``synthdata`` (pure data generators: weights with
the reference's state-dict key names, frames, analytic scenes -- no algorithm of the path).  Re-exported here so
the oracle and the tests keep one name for them."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)

from synthdata import *  # noqa: F401,F403,E402
from synthdata import LFN_BACKWARD, LFN_DIST_CH, LFN_FEAT_CH, LFN_KLAST, LFN_LEVELS, LFN_REG_CIN, LFN_SUB_CIN  # noqa: F401,E402
