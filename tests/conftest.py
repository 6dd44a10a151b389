"""pytest configuration: markers and import paths.

``-m "not gpu"`` runs the oracle-vs-golden, host-logic and C-ABI symbol tests on CPU;
``-m gpu`` runs the parity tests proper on a B200 through the C-ABI library.
"""

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semantic-slam-master_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
