"""Flat module name imported by src/training/training.py:31 (`from complete_model import ...`; the file does not
exist in the reference)."""
import _path  # noqa: F401
from deer_b200.complete_project import *  # noqa: F401,F403
from deer_b200.complete_project import CompleteDEERModel, ModelCheckpoint, ModelConfig  # noqa: F401
