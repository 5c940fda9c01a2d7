"""Flat module name imported by run_multimodal_deer.py:83.  Plotting is out of scope (SURVEY.md section 2 #9;
matplotlib/seaborn are not in this image): the report is written as JSON next to where the plots would go."""
import json
import os

import numpy as np


def create_comprehensive_report(predictions, targets, uncertainties=None, training_history=None, save_dir="./plots",
                                report_name="deer_report", **_):
    from metrics import DEERMetrics
    os.makedirs(save_dir, exist_ok=True)
    rep = {"n_samples": int(np.asarray(predictions).shape[0]),
           "metrics": DEERMetrics().compute_all_metrics(targets, predictions, uncertainties),
           "training_history_keys": sorted((training_history or {}).keys())}
    path = os.path.join(save_dir, f"{report_name}.json")
    with open(path, "w") as f:
        json.dump(rep, f, indent=2)
    return path


def test_visualization_components():
    return True
