"""deer_b200: hand-written sm_100a CUDA (libdeer_b200.so, C ABI) behind the reference's module/loss API for the
batched forward+backward of the multimodal DEER model.  See DESIGN.md / INTEGRATION.md."""
from . import _lib, ops  # noqa: F401
from .complete_project import CompleteDEERModel, ModelCheckpoint, ModelConfig, create_complete_deer_model  # noqa: F401
from .deer import DEERLayer, DEERLoss as AminiDEERLoss, MultiDimensionalDEER  # noqa: F401
from .encoders import (AudioEncoder, EnhancedAudioEncoder, EnhancedTextEncoder, EnhancedVideoEncoder,  # noqa: F401
                       TextEncoder, VideoEncoder)
from .fusion import AudioVisualFusion, HierarchicalMultimodalFusion, TrimodalFusion  # noqa: F401
from .losses import CombinedDEERLoss, DEERLoss, MultiTaskDEERLoss, create_deer_loss  # noqa: F401
from .sequence_model import SequenceDEERModel  # noqa: F401

__version__ = "0.1.0"
