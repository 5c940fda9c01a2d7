"""Import alias: `import deer_b200` -> the package directory `uncertainty-aware-multimodal-emotion-recognition_b200`
(whose name is not a valid Python identifier)."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_pkg = importlib.import_module("uncertainty-aware-multimodal-emotion-recognition_b200")
sys.modules[__name__] = _pkg
