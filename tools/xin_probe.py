"""Times the first LSTM layer's forward (B x T x 84 -> [T,B,512]) with the input projection inside the recurrence kernel
(deer_lstm_cluster_fwd_xin) and as a GEMM + FP16 pre-activations; train (kept state) and eval."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deer_b200  # noqa: E402,F401
from deer_b200 import ops  # noqa: E402
from deer_b200.encoders import EnhancedAudioEncoder  # noqa: E402

enc = EnhancedAudioEncoder({"dropout": 0.0}).cuda()
w = enc._layer_weights(0)


def run(B, T, train):
    x = torch.randn(B, T, 84, device="cuda")
    enc.train(train)
    out = {}
    for fused in (1, 0, 1, 0):
        ops.set_lstm_input_projection_fused(bool(fused))
        def f():
            if train:
                return ops.bilstm_layer(x, *w, x_batch_major=True, return_bf16=True)
            with torch.no_grad():
                return ops.bilstm_layer(x, *w, x_batch_major=True, return_bf16=True)
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            f()
        e1.record()
        torch.cuda.synchronize()
        out.setdefault(fused, []).append(e0.elapsed_time(e1) / 10 * 1e3)
    print(f"B={B} T={T} train={train}: layer-0 forward (prep + cast + [GEMM] + recurrence)  in-kernel projection "
          f"{min(out[1]):7.1f} us | GEMM path {min(out[0]):7.1f} us", flush=True)


run(256, 300, True)
run(256, 300, False)
run(1024, 300, False)
ops.set_lstm_input_projection_fused(True)
