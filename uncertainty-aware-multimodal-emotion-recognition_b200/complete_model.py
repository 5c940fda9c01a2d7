"""Alias module: the reference trainer imports `complete_model` (src/training/training.py:31) although the code
lives in complete_project.py (file-name trap)."""
from .complete_project import *  # noqa: F401,F403
from .complete_project import CompleteDEERModel, ModelCheckpoint, ModelConfig, create_complete_deer_model  # noqa: F401
