"""Flat module name imported by run_multimodal_deer.py:72 (`MultiDatasetDEERFramework`; the reference class is
actually called MultiDatasetFramework, multi_dataset_framework.py:361, and its "training" returns constants,
:446-455).  Only the batch wire format of its samples matters to the hot path (:85-103): 84/256/768-D vectors,
targets[3], dataset_id, with per-dataset loss weights 1.0/0.8/0.6 (training.py:59-61)."""
import torch

DATASETS = ("IEMOCAP", "RAVDESS", "MELD")
DATASET_LOSS_WEIGHTS = {"IEMOCAP": 1.0, "RAVDESS": 0.8, "MELD": 0.6}


class MultiDatasetFramework:
    def __init__(self, config=None):
        self.config = config or {}

    @staticmethod
    def synthetic_batch(batch_size: int, device="cpu", seed: int = 0):
        g = torch.Generator().manual_seed(seed)
        b = {"audio_features": torch.randn(batch_size, 84, generator=g),
             "video_features": torch.randn(batch_size, 256, generator=g),
             "text_features": torch.randn(batch_size, 768, generator=g),
             "targets": torch.tanh(torch.randn(batch_size, 3, generator=g)),
             "dataset_id": torch.randint(0, len(DATASETS), (batch_size, 1), generator=g)}
        return {k: v.to(device) for k, v in b.items()}


MultiDatasetDEERFramework = MultiDatasetFramework
