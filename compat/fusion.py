"""Flat module name imported by run_multimodal_deer.py:79."""
import _path  # noqa: F401
from deer_b200.fusion import *  # noqa: F401,F403
from deer_b200.fusion import AudioVisualFusion, HierarchicalMultimodalFusion, TrimodalFusion  # noqa: F401
