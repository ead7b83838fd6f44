"""Drop-in for the reference's src/models/generator.py: same class names, backed by the sm_100a kernels."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from pbt_b200.generator import GeneratorJ, ResNetBlock, UpsamplingLayer  # noqa: E402,F401
