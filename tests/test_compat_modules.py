"""The 11 flat module names the unchanged reference driver imports (experiments/run_multimodal_deer.py:72-82, SURVEY.md
section 8b) resolve from `compat/` and export the names it asks for; plus the GPU run of the driver's call sequence."""
import importlib
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "compat")

DRIVER_IMPORTS = {
    "multi_dataset_framework": ["MultiDatasetDEERFramework"],
    "complete_project": ["CompleteDEERModel", "ModelConfig"],
    "training": ["DEERTrainer", "TrainingConfig"],
    "preprocessing": ["create_enhanced_dataloaders"],
    "deer": ["test_deer_implementation"],
    "encoders": ["AudioEncoder", "VideoEncoder", "TextEncoder"],
    "fusion": ["HierarchicalMultimodalFusion"],
    "evaluation": ["evaluate_deer_model"],
    "metrics": ["DEERMetrics"],
    "losses": ["DEERLoss"],
    "visualization": ["create_comprehensive_report", "test_visualization_components"],
    "complete_model": ["CompleteDEERModel", "ModelConfig", "ModelCheckpoint"],   # src/training/training.py:31
}


@pytest.fixture()
def compat_path():
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in DRIVER_IMPORTS}
    sys.path.insert(0, COMPAT)
    for k in DRIVER_IMPORTS:
        sys.modules.pop(k, None)
    yield
    sys.path[:] = saved_path
    for k, v in saved_mods.items():
        sys.modules.pop(k, None)
        if v is not None:
            sys.modules[k] = v


def test_flat_module_names_resolve(compat_path):
    for mod, names in DRIVER_IMPORTS.items():
        m = importlib.import_module(mod)
        assert os.path.dirname(os.path.abspath(m.__file__)) == COMPAT, (mod, m.__file__)
        for n in names:
            assert hasattr(m, n), (mod, n)


@pytest.mark.gpu
def test_compat_metrics_take_numpy(compat_path):
    """`from metrics import DEERMetrics` (run_multimodal_deer.py:81): NumPy in, device-side reductions."""
    metrics = importlib.import_module("metrics")
    m = metrics.DEERMetrics()
    x = np.linspace(-1, 1, 50)
    assert abs(m.concordance_correlation_coefficient(x, x) - 1.0) < 1e-6
    assert abs(m.concordance_correlation_coefficient(x, -x) + 1.0) < 1e-6
    assert m.concordance_correlation_coefficient(x, x + 0.5) < 1.0
    r = m.evaluate_predictions(np.stack([x, x, x], 1), np.stack([x, -x, x + 0.5], 1), np.abs(np.stack([x, x, x], 1)))
    assert abs(r.ccc_valence - 1.0) < 1e-6 and abs(r.ccc_arousal + 1.0) < 1e-6 and 0 <= r.ece <= 2


def test_host_loaders(compat_path):
    pre = importlib.import_module("preprocessing")
    tr, va, te = pre.create_enhanced_dataloaders(config={"model": {"audio_dim": 84, "video_dim": 256, "text_dim": 768}},
                                                 batch_size=8)
    assert [len(next(iter(d.values())).dataset) for d in (tr, va, te)] == [1000, 200, 200]
    batch = next(iter(next(iter(te.values()))))
    assert len(batch) == 4 and batch[0].shape == (8, 84) and batch[3].shape == (8, 3)
    assert float(batch[3].abs().max()) <= 1.0


@pytest.mark.gpu
def test_driver_call_sequence_on_gpu(compat_path, tmp_path):
    """The calls MultimodalDEERPipeline.run_full_pipeline makes (run_multimodal_deer.py:231-760) with --quick sizes
    (5 epochs, batch 8 in the reference; 2 epochs here), everything json.dump-able as the driver requires."""
    cp = importlib.import_module("complete_project")
    training = importlib.import_module("training")
    pre = importlib.import_module("preprocessing")
    evaluation = importlib.import_module("evaluation")
    viz = importlib.import_module("visualization")
    device = torch.device("cuda")
    torch.manual_seed(42)
    cfg = cp.ModelConfig(audio_dim=84, video_dim=256, text_dim=768, fusion_dim=512, emotion_dims=3, dropout=0.3,
                         attention_heads=8)
    model = cp.CompleteDEERModel(cfg).to(device)
    assert sum(p.numel() for p in model.parameters()) == 3918324
    tr, va, te = pre.create_enhanced_dataloaders(config={"model": {}}, batch_size=8, sizes=(96, 32, 32))
    tcfg = training.TrainingConfig(learning_rate=1e-4, batch_size=8, num_epochs=2, weight_decay=1e-5, gradient_clip=1.0,
                                   output_dir=str(tmp_path / "models"), log_dir=str(tmp_path / "logs"))
    trainer = training.DEERTrainer(model, tcfg, device)
    history = trainer.train(tr, va)
    json.dumps(history)
    assert len(history["train_loss"]) == 2 and all(np.isfinite(history["train_loss"]))
    assert history["val_ccc"] and history["val_loss"]
    results = evaluation.evaluate_deer_model(model, te, device=device, save_predictions=True, save_dir=str(tmp_path))
    json.dumps(results)
    assert os.path.exists(tmp_path / "predictions.npz") and results["n_samples"] == 32
    # visualisation sampling loop (:696-729): dict input, 'gamma' in outputs, CPU targets
    model.eval()
    with torch.no_grad():
        audio, video, text, emotions = next(iter(next(iter(te.values()))))
        outputs = model({"audio": audio.to(device), "video": video.to(device), "text": text.to(device)})
        assert isinstance(outputs, dict) and "gamma" in outputs
        preds, uncs = model.get_predictions_and_uncertainties(outputs)
    path = viz.create_comprehensive_report(predictions=preds.cpu().numpy(), targets=emotions.numpy(),
                                           uncertainties=uncs.cpu().numpy(), training_history=history,
                                           save_dir=str(tmp_path / "plots"), report_name="t")
    assert os.path.exists(path)
    # checkpoint round trip with reference-named keys (:512-517, :931-933)
    torch.save({"model_state_dict": model.state_dict(), "training_history": history}, tmp_path / "final_model.pth")
    model2 = cp.CompleteDEERModel(cfg).to(device)
    model2.load_state_dict(torch.load(tmp_path / "final_model.pth", weights_only=False)["model_state_dict"])
