"""Flat module name imported by run_multimodal_deer.py:82 (`from losses import DEERLoss`)."""
import _path  # noqa: F401
from deer_b200.losses import *  # noqa: F401,F403
from deer_b200.losses import CombinedDEERLoss, DEERLoss, MultiTaskDEERLoss, create_deer_loss  # noqa: F401
