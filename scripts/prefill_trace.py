#!/usr/bin/env python
"""Tuning aid (GPU box): run the C3 prefill once on a library built with -DTVMB200_TRACE=<block> and print the
clock64 stamps of that CTA: per KV step, how long the MMA warp waited for P / took to issue, and how long the
softmax warps waited for S / computed.  usage: TVMB200_LIB_SUFFIX=_trace python scripts/prefill_trace.py [dtype]"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from tvm_b200 import capi  # noqa: E402

dt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
w = bench.PrefillWorkload(dtype=dt) if hasattr(bench, "PrefillWorkload") else None
if w is None:
    cls = [v for k, v in vars(bench).items() if isinstance(v, type) and hasattr(v, "flops") and "L" in v.__init__.__code__.co_varnames]
    w = cls[0](dtype=dt)
for _ in range(3):
    w.run(capi)
torch.cuda.synchronize()
L = capi.lib()
buf = (ctypes.c_longlong * (3 * 64 * 8))()
rc = L.tvmb200_debug_prefill_trace(buf)
assert rc == 0, rc
tr = np.frombuffer(buf, dtype=np.int64).reshape(3, 64, 8)
t0 = tr[tr > 0].min()
rel = np.where(tr > 0, tr - t0, -1)
mode = sys.argv[2] if len(sys.argv) > 2 else "tile"
if mode == "tile":
    # softmax warpgroup t owns Q tile t; stamps indexed by half-step s = 2j+h
    print("role 1+t = softmax warpgroup of Q tile t, index s: [0] wait S, [1] S ready, [2] ld done, [3] P arrive")
    print("role 0 = MMA warp, index s: [3t+1] P ready seen, [3t+2] PV(+QK) issued.   times in clk from the first stamp")
    print("  s | T0: Swait   ld  comp  arrive@ | T1: Swait   ld  comp  arrive@ | MMA t0: seen@ issued@ (dur) | t1: seen@ issued@ (dur)")
    for s in range(64):
        a, b, m = rel[1, s], rel[2, s], rel[0, s]
        if a[3] < 0 and b[3] < 0:
            break
        print(f"{s:3d} | {a[1]-a[0]:6d} {a[2]-a[1]:5d} {a[3]-a[2]:5d} {a[3]:8d} | {b[1]-b[0]:6d} {b[2]-b[1]:5d} {b[3]-b[2]:5d} {b[3]:8d} |"
              f" {m[1]:7d} {m[2]:7d} ({m[2]-m[1]:4d}) | {m[4]:7d} {m[5]:7d} ({m[5]-m[4]:4d})"
              f" || T0 phases: max {a[4]-a[2]:4d} exp {a[5]-a[4]:4d} st+lo {a[6]-a[5]:4d} wait_st+arrive {a[3]-a[6]:4d}"
              f" | T1: max {b[4]-b[2]:4d} exp {b[5]-b[4]:4d} st+lo {b[6]-b[5]:4d} wait_st+arrive {b[3]-b[6]:4d}")
else:
    print("role 0 (MMA warp) index s = 2j+h: [3t+1] P ready seen, [3t+2] PV(+QK) issued;  roles 1/2 (softmax warpgroup 0/1)")
    print("index 2j+t: [0] wait S, [1] S ready, [2] ld done, [3] P arrive.   all times in clk from the first stamp")
    print(" j t | WG0: Swait   ld  comp  arrive@ | WG1: Swait   ld  comp  arrive@ | MMA h0: seen@ issued@ (dur) | h1: seen@ issued@ (dur)")
    for j in range(32):
        for tt in range(2):
            a, b = rel[1, 2 * j + tt], rel[2, 2 * j + tt]
            if a[3] < 0:
                continue
            m0, m1 = rel[0, 2 * j], rel[0, 2 * j + 1]
            print(f"{j:2d} {tt} | {a[1]-a[0]:6d} {a[2]-a[1]:5d} {a[3]-a[2]:5d} {a[3]:8d} | {b[1]-b[0]:6d} {b[2]-b[1]:5d} {b[3]-b[2]:5d} {b[3]:8d} |"
                  f" {m0[3*tt+1]:7d} {m0[3*tt+2]:7d} ({m0[3*tt+2]-m0[3*tt+1]:4d}) | {m1[3*tt+1]:7d} {m1[3*tt+2]:7d} ({m1[3*tt+2]-m1[3*tt+1]:4d})")
