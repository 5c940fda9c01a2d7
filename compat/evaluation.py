"""Flat module name imported by run_multimodal_deer.py:80.  The driver calls
`evaluate_deer_model(model, test_loaders, device=, save_predictions=True, save_dir=)` (:532-538) and json.dumps the
result; the reference's own signature differs (evaluation.py:785).  Forward passes run on the
CUDA path; the statistics are device-side reductions (deer_b200.metrics)."""
import os

import numpy as np
import torch

from metrics import DEERMetrics


def _unpack(batch, device):
    if isinstance(batch, dict):
        a, v, t, y = (batch[k] for k in ("audio_features", "video_features", "text_features", "targets"))
    else:
        a, v, t, y = batch
    return [x.to(device, dtype=torch.float32, non_blocking=True) for x in (a, v, t)], y


@torch.no_grad()
def collect_predictions(model, loaders, device):
    was_training = model.training
    model.eval()
    P, U, Y = [], [], []
    loaders = loaders.values() if isinstance(loaders, dict) else [loaders]
    for loader in loaders:
        for batch in loader:
            (a, v, t), y = _unpack(batch, device)
            out = model(a, v, t)
            p, u = model.get_predictions_and_uncertainties(out)
            P.append(p.float())
            U.append(u.float())
            Y.append(y.to(device, dtype=torch.float32, non_blocking=True))
    model.train(was_training)
    return torch.cat(P), torch.cat(U), torch.cat(Y)  # stay on the device: the metric reductions are kernels


def evaluate_deer_model(model, dataloader, device=None, config=None, save_predictions: bool = False, save_dir=None):
    device = device or next(model.parameters()).device
    preds, uncs, tgts = collect_predictions(model, dataloader, device)
    results = DEERMetrics().compute_all_metrics(tgts, preds, uncs)
    results["n_samples"] = int(preds.shape[0])
    if save_predictions and save_dir:
        os.makedirs(save_dir, exist_ok=True)
        np.savez(os.path.join(save_dir, "predictions.npz"), predictions=preds.cpu().numpy(),
                 uncertainties=uncs.cpu().numpy(), targets=tgts.cpu().numpy())
    return results
