"""Flat module name imported by run_multimodal_deer.py:77 (`from deer import test_deer_implementation`)."""
import _path  # noqa: F401
from deer_b200.deer import *  # noqa: F401,F403
from deer_b200.deer import DEERLayer, DEERLoss, MultiDimensionalDEER, test_deer_implementation  # noqa: F401
