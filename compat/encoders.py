"""Flat module name imported by run_multimodal_deer.py:78 (`from encoders import AudioEncoder, VideoEncoder, TextEncoder`)."""
import _path  # noqa: F401
from deer_b200.encoders import *  # noqa: F401,F403
from deer_b200.encoders import (AudioEncoder, EnhancedAudioEncoder, EnhancedTextEncoder,  # noqa: F401
                                EnhancedVideoEncoder, TextEncoder, VideoEncoder)
