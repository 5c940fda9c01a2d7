"""Flat module name imported by run_multimodal_deer.py:81 (`from metrics import DEERMetrics`).  The reductions run on
the device (deer_b200.metrics: moments kernel for CCC / MAE / RMSE / Cohen's d, radix-select + binning kernels for the
quantile calibration error, src/utils/metrics.py:59-279); NumPy inputs, as the reference API takes them, are uploaded
first."""
import numpy as np
import torch

import _path  # noqa: F401
from deer_b200.metrics import (EvaluationResults, DEERMetrics as _DeviceMetrics,  # noqa: F401
                               uncertainty_calibration_error as _device_uce)


def _dev(x):
    if x is None or (torch.is_tensor(x) and x.is_cuda):
        return x
    return torch.as_tensor(np.asarray(x), dtype=torch.float32).cuda()


class DEERMetrics(_DeviceMetrics):
    def concordance_correlation_coefficient(self, y_true, y_pred) -> float:
        return super().concordance_correlation_coefficient(_dev(y_true), _dev(y_pred))

    def mean_absolute_error(self, y_true, y_pred) -> float:
        return super().mean_absolute_error(_dev(y_true), _dev(y_pred))

    def root_mean_squared_error(self, y_true, y_pred) -> float:
        return super().root_mean_squared_error(_dev(y_true), _dev(y_pred))

    def evaluate_predictions(self, predictions, targets, uncertainties=None):
        return super().evaluate_predictions(_dev(predictions), _dev(targets), _dev(uncertainties))

    def compute_all_metrics(self, y_true, y_pred, uncertainties=None):
        return super().compute_all_metrics(_dev(y_true), _dev(y_pred), _dev(uncertainties))


def uncertainty_calibration_error(predictions, targets, uncertainties, n_bins: int = 10) -> float:
    return _device_uce(_dev(predictions), _dev(targets), _dev(uncertainties), n_bins)
