"""Drop-in for the reference's src/data/dataset.py: StyleTransferDataset with device-resident keyframes."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from pbt_b200.sampler import StyleTransferDataset  # noqa: E402,F401
