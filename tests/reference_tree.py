"""TEST INFRASTRUCTURE — the loaders of the UNMODIFIED reference live in baseline/reference_loader.py
(bench.py's reference arm uses them too); re-exported here for the tests."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline"))
from reference_loader import adapter, render_adapter, stock, tree_root  # noqa: E402,F401
