"""Flat module name imported by run_multimodal_deer.py:74 (`from training import DEERTrainer, TrainingConfig`).

Mirrors the trainer-facing contract of src/training/training.py (TrainingConfig :39-72, DEERTrainer :75-535):
AdamW with the 0.5x LR group for parameters whose name contains "encoder" (:121-150), cosine LR with eta_min 1e-6
(:152-174), gradient clipping (:219), per-epoch history.  The step itself is deer_b200's fused trainer step (flat
buffers, NCCL all-reduce when launched under torch.distributed, fused clip+AdamW kernel); accepts the reference's dict
batches (:201-204) and the driver's 4-tuple synthetic batches (run_multimodal_deer.py:342)."""
import math
import time
from dataclasses import dataclass
from typing import Dict, List

import torch

import _path  # noqa: F401
from deer_b200.trainer import DEERDataParallelTrainer
from evaluation import collect_predictions
from metrics import DEERMetrics


@dataclass
class TrainingConfig:
    learning_rate: float = 1e-4
    batch_size: int = 32
    num_epochs: int = 100
    weight_decay: float = 1e-5
    gradient_clip: float = 1.0
    scheduler_type: str = "cosine"
    min_lr: float = 1e-6
    validation_frequency: int = 5
    early_stopping_patience: int = 20
    output_dir: str = "./outputs"
    log_dir: str = "./logs"
    use_cuda_graph: bool = True   # replay the captured step (one launch) instead of ~300 kernel launches per batch


class DEERTrainer:
    def __init__(self, model, config: TrainingConfig, device=None):
        self.config = config
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.model = model.to(self.device)
        self.step = DEERDataParallelTrainer(self.model, learning_rate=config.learning_rate,
                                            weight_decay=config.weight_decay, gradient_clip=config.gradient_clip)
        self.metrics = DEERMetrics()
        self.history: Dict[str, List[float]] = {"train_loss": [], "val_loss": [], "val_ccc": [], "learning_rate": [],
                                                "epoch_time": []}

    def _lr(self, epoch: int) -> float:
        c = self.config
        if c.scheduler_type != "cosine" or c.num_epochs <= 1:
            return c.learning_rate
        return c.min_lr + 0.5 * (c.learning_rate - c.min_lr) * (1.0 + math.cos(math.pi * epoch / c.num_epochs))

    def _batch(self, batch) -> Dict[str, torch.Tensor]:
        if not isinstance(batch, dict):
            a, v, t, y = batch
            batch = {"audio_features": a, "video_features": v, "text_features": t, "targets": y}
        return {k: x.to(self.device, dtype=torch.float32, non_blocking=True) for k, x in batch.items()
                if torch.is_tensor(x) and k != "dataset_id"}

    def train_epoch(self, train_loaders) -> float:
        self.model.train()
        total = torch.zeros((), device=self.device)
        n = 0
        for loader in train_loaders.values():
            for batch in loader:
                b = self._batch(batch)
                losses = self.step.train_step_auto(b) if self.config.use_cuda_graph else self.step.train_step(b)
                total += losses[-1]            # stays on the device: no per-step host sync (cf. training.py:228-230)
                n += 1
        return float(total) / max(n, 1)

    @torch.no_grad()
    def validate_epoch(self, val_loaders) -> Dict[str, float]:
        preds, uncs, tgts = collect_predictions(self.model, val_loaders, self.device)
        m = self.metrics.compute_all_metrics(tgts, preds, uncs)
        self.model.eval()
        total, n = 0.0, 0
        for loader in val_loaders.values():
            for batch in loader:
                b = self._batch(batch)
                total += float(self.model.compute_loss(self.model(b), b["targets"])["total_loss"])
                n += 1
        self.model.train()
        return {"val_loss": total / max(n, 1), "val_ccc": m["mean_ccc"], **m}

    def train(self, train_loaders, val_loaders=None) -> Dict[str, List[float]]:
        c = self.config
        for epoch in range(c.num_epochs):
            t0 = time.time()
            self.step.lr = self._lr(epoch)
            self.history["train_loss"].append(self.train_epoch(train_loaders))
            self.history["learning_rate"].append(self.step.lr)
            if val_loaders and ((epoch + 1) % c.validation_frequency == 0 or epoch == c.num_epochs - 1):
                v = self.validate_epoch(val_loaders)
                self.history["val_loss"].append(v["val_loss"])
                self.history["val_ccc"].append(v["val_ccc"])
            self.history["epoch_time"].append(time.time() - t0)
        return self.history

    def evaluate_model(self, test_loaders) -> Dict[str, float]:
        preds, uncs, tgts = collect_predictions(self.model, test_loaders, self.device)
        return self.metrics.compute_all_metrics(tgts, preds, uncs)
