"""Flat module name imported by experiments/run_multimodal_deer.py:73 (`from complete_project import ...`)."""
import _path  # noqa: F401
from deer_b200.complete_project import *  # noqa: F401,F403
from deer_b200.complete_project import (CompleteDEERModel, ModelCheckpoint, ModelConfig,  # noqa: F401
                                         create_complete_deer_model)
