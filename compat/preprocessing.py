"""Flat module name imported by run_multimodal_deer.py:75.  The reference's IEMOCAP loader needs dataset files,
librosa and BERT (src/data/preprocessing.py:57-787, out of scope); the driver calls
`create_enhanced_dataloaders(config=<dict>, batch_size=<int>)` and expects three dicts name -> DataLoader
(:317-320).  This provides exactly the synthetic loaders the driver itself falls back to (:329-349): randn features of
the configured widths, targets tanh(randn + 0.1 randn), N = 1000/200/200, 4-tuples per batch."""
import torch
from torch.utils.data import DataLoader, TensorDataset


def create_enhanced_dataloaders(config=None, batch_size: int = 32, root_path=None, num_workers: int = 0, seed: int = 42,
                                sizes=(1000, 200, 200), **_):
    model_cfg = (config or {}).get("model", {}) if isinstance(config, dict) else {}
    dims = (model_cfg.get("audio_dim", 84), model_cfg.get("video_dim", 256), model_cfg.get("text_dim", 768))
    g = torch.Generator().manual_seed(seed)
    out = []
    for n, split in zip(sizes, ("train", "val", "test")):
        feats = [torch.randn(n, d, generator=g) for d in dims]
        emotions = torch.tanh(torch.randn(n, 3, generator=g) + 0.1 * torch.randn(n, 3, generator=g))
        loader = DataLoader(TensorDataset(*feats, emotions), batch_size=batch_size, shuffle=(split == "train"),
                            generator=torch.Generator().manual_seed(seed + 1), pin_memory=torch.cuda.is_available())
        out.append({f"synthetic_{split}": loader})
    return tuple(out)
