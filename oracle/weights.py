"""Re-export of the synthetic fixture generators (synth/weights.py) for the oracle-side scripts and tests."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synth.weights import *  # noqa: F401,F403,E402
from synth.weights import _normal, _rng, _uniform  # noqa: F401,E402
