"""Makes the ``paresis_b200`` package importable from this CodePython-shaped directory
(the reference is run with cwd = CodePython and flat imports, main.py:12-15)."""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
