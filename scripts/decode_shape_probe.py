#!/usr/bin/env python
"""tuning helper (GPU box): the fused decode step at several (B, L, Hq, Hkv) shapes with the same KV bytes per GPU --
C2, and the per-rank shards of C4 at tp = 2 / 4 / 8 -- GB/s of the step's algorithmic bytes."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import bench  # noqa: E402
from tvm_b200 import capi  # noqa: E402

dev = torch.device("cuda", 0)
capi.lib()
for name, (B, L, Hq, Hkv) in {"C2": (64, 4096, 32, 8), "C4/tp8": (256, 8192, 8, 1), "C4/tp4": (256, 8192, 16, 2),
                              "C4/tp2": (256, 8192, 32, 4), "C2 g8": (64, 4096, 64, 8), "tp8 B64 L32k": (64, 32768, 8, 1)}.items():
    w = bench.DecodeWorkload(B=B, L=L, Hq=Hq, Hkv=Hkv, device=dev)
    for _ in range(5):
        w.run_step_fused(capi)
    ms, win, _ = bench.time_windows(lambda: w.run_step_fused(capi), 50, 5, dev)
    print(f"{name:14s} B {B} L {L} Hq {Hq} Hkv {Hkv}: {ms*1e3:.1f} us  {w.step_bytes()/ms/1e6:.0f} GB/s  windows {win}")
    del w
    bench._free()
