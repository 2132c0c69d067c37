#!/usr/bin/env python
"""bench.py -- PagedKVCache attention hot path on B200 (contract: see DESIGN.md "Measurement").

Headline (the JSON line's `value`): one layer-call of the decode hot path over one batch, exactly the callback sequence the
reference cache issues per layer for a plain decode step (paged_kv_cache.cc:1355-1401):
    f_split_rotary -> f_transpose_append -> f_attention_decode
on BASELINE.json configs[1] (C2: Llama-3-8B shape, batch 64 decode at 4K context, bf16 paged KV).  With --gpus N the
sequence batch is split across the ranks (64 whole sequences per rank, weak scaling): a sequence's attention output feeds
that rank's own next layer, so this mode has NO data-path collective; NCCL carries the barrier and the max-over-ranks only.

    metric  decode_attn_hbm_gbps = algorithmic bytes of the step (BASELINE.md section 3 formulas) / time
    value   device-resident inputs; R windows of exactly K steps, CUDA events, max over ranks per window, median window
    e2e     the same through the public host-cache API with HOST inputs (pinned qkv in, O out, aux arrays H2D), every step
    roofline  the step's launch pair (fused decode + split-KV merge) against the measured HBM copy peak
    cpu_baseline  the reference's CPU path (oracle/_ref if built, else the NumPy oracle port) on a slice

The rest of BASELINE.json's metric rides in the same line as sub-records (each with its own windows / clocks / roofline):
    prefill_c3        C3 ragged causal prefill 16 x 2048 (TFLOP/s, tcgen05 kernel)
    append_c3         C3's KV append, 32768 tokens: f_transpose_append alone and the fused split_rotary+append (GB/s)
    c5_tree_prefill   C5: 64-node token trees x batch 32 over a 32K-token cached context (tree-masked self part +
                      mask-free paged part + merge, the callback sequence of the first tree round)
    c5_decode_32k     C5: split-KV decode, batch 32 at 32K context (GB/s)
    ref_gpu           the reference's OWN GPU TIR kernels (oracle/_ref, a subprocess) on the C2 / C3 inputs: kernel to beat
    c4_head_sharded   N > 1 only: C4 (70B GQA decode, batch 256 x 8K) with KV-head groups sharded tp = N (strong scaling)
                      and the per-head outputs re-assembled on every rank (in-kernel NVLink peer gather / NCCL all-gather)

`--impl reference` times the CPU path only (rank 0) and prints the same JSON line shape.
`--workload prefill|append|c5|c5decode|c4` print one sub-record as the line (profiler runs); `--no-sub` skips them.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of the dominant kernel on the named workload,
# read from the committed `ncu --set full` summaries under profiles/ at run time (a profiler number: reported next to
# the algorithmic bytes, never used for timing).  profiles/ncu_traffic.json: {workload: {"bytes": .., "source": ..}}
NCU_TRAFFIC_FILE = ROOT / "profiles" / "ncu_traffic.json"


def ncu_traffic(key):
    try:
        d = json.loads(NCU_TRAFFIC_FILE.read_text())[key]
        return int(d["bytes"]), d["source"]
    except Exception:
        return None, None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return d, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


# ---------------------------------------------------------------------------------------------------
# algorithmic work: pure functions, checked against the table's numbers in tests/test_bench_contract.py
# ---------------------------------------------------------------------------------------------------
def decode_bytes_of(B, L, Hq, Hkv, D, page=16, e=2):
    """f_attention_decode: every KV byte once, q in, O / LSE out, the index arrays"""
    nnz = B * (-(-L // page))
    return B * L * Hkv * D * 2 * e + 2 * B * Hq * D * e + 4 * B * Hq + 4 * (nnz + 4 * B + 1)


def append_bytes_of(n, Hkv, D, e=2):
    """f_transpose_append: k, v read and written into their slots, the slot ids"""
    return n * Hkv * D * 2 * e * 2 + 4 * n


def causal_prefill_flops_of(nseq, L, Hq, D):
    """4 * D * Hq * (# unmasked (q, k) pairs), empty cache"""
    return 4 * D * Hq * nseq * (L * (L + 1) // 2)


# ---------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------
class DecodeWorkload:
    """B sequences of L cached tokens (after this step's append), Hq/Hkv heads, D=128, page 16, bf16."""

    def __init__(self, B=64, L=4096, Hq=32, Hkv=8, D=128, page=16, seed=0, device="cuda"):
        import torch

        self.B, self.L, self.Hq, self.Hkv, self.D, self.page = B, L, Hq, Hkv, D, page
        rng = np.random.default_rng(seed)
        ppseq = -(-L // page)
        self.nnz = B * ppseq
        P = self.nnz + 1
        self.P = P
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        self.pages = torch.randn((P, 2, Hkv, page, D), generator=g, device=device, dtype=torch.bfloat16)
        perm = rng.permutation(P).astype(np.int32)[: self.nnz]  # non-contiguous page ids: a real gather
        self.h_page_values = perm
        self.h_page_indptr = (np.arange(B + 1) * ppseq).astype(np.int32)
        self.h_length_info = np.full(B, ((L - 1) % page) + 1, np.int32)
        self.h_k_rope_pos_offset = np.zeros(B, np.int32)
        self.h_q_rope_position = np.full(B, L - 1, np.int32)
        last_page = perm.reshape(B, ppseq)[:, -1]
        self.h_append_position = (last_page * page + (L - 1) % page).astype(np.int32)
        self.h_qkv = torch.randn((B, Hq + 2 * Hkv, D), generator=torch.Generator().manual_seed(seed),
                                 dtype=torch.float32).to(torch.bfloat16).pin_memory()
        i32 = lambda a: torch.from_numpy(a).to(device)  # noqa: E731
        self.page_values, self.page_indptr = i32(self.h_page_values), i32(self.h_page_indptr)
        self.length_info, self.k_rope_pos_offset = i32(self.h_length_info), i32(self.h_k_rope_pos_offset)
        self.q_rope_position, self.append_position = i32(self.h_q_rope_position), i32(self.h_append_position)
        self.qkv = self.h_qkv.to(device)
        self.q = torch.empty((B, Hq, D), device=device, dtype=torch.bfloat16)
        self.k = torch.empty((B, Hkv, D), device=device, dtype=torch.bfloat16)
        self.v = torch.empty((B, Hkv, D), device=device, dtype=torch.bfloat16)
        self.o = torch.empty((B, Hq, D), device=device, dtype=torch.bfloat16)
        self.lse = torch.empty((B, Hq), device=device, dtype=torch.float32)
        self.sm_scale = D ** -0.5
        self.rope_theta, self.rope_scale = 5e5, 1.0

    # algorithmic bytes
    def decode_bytes(self):
        return decode_bytes_of(self.B, self.L, self.Hq, self.Hkv, self.D, self.page)

    def append_bytes(self):
        return append_bytes_of(self.B, self.Hkv, self.D)

    def step_bytes(self):
        """the fused step: f_attention_decode's bytes (q in, O / LSE out, every KV byte once, the index arrays) + the
        append's (new k, v read from the fused qkv and written into their page slots, the slot ids) + the rotary's one
        array the others do not already count (q_rope_position).  The q / k / v round trip through HBM that a separate
        f_split_rotary launch performs does not exist in the fused launch and is NOT counted."""
        return self.decode_bytes() + self.append_bytes() + 4 * self.B

    def run_rotary_append(self, capi):
        capi.split_rotary_append(self.qkv, self.q_rope_position, self.append_position, self.q, self.k, self.v,
                                 self.pages, 1, self.rope_scale, self.rope_theta)

    def run_step_fused(self, capi):
        """the whole step -- f_split_rotary + f_transpose_append + f_attention_decode -- as one launch (+ the merge)"""
        capi.attention_decode_fused_qkv(self.qkv, self.q_rope_position, self.append_position, self.pages, self.page_indptr,
                                        self.page_values, self.length_info, self.k_rope_pos_offset, self.o, self.lse, 1,
                                        self.rope_scale, self.rope_theta, self.sm_scale)

    def run_decode_gather(self, capi, gather):
        """decode of this rank's KV-head shard; the kernel stores its heads into every rank's gathered buffer"""
        return gather.decode(capi, self.q, self.pages, self.page_indptr, self.page_values, self.length_info,
                             self.k_rope_pos_offset, self.q_rope_position, self.o, self.lse, 0, self.rope_scale,
                             self.rope_theta, self.sm_scale)

    def run_decode(self, capi):
        capi.attention_decode(self.q, self.pages, self.page_indptr, self.page_values, self.length_info,
                              self.k_rope_pos_offset, self.q_rope_position, self.o, self.lse, 0, self.rope_scale,
                              self.rope_theta, self.sm_scale)


class PrefillWorkload:
    """C3: 16 sequences x 2048 new tokens, empty cache -> ragged causal prefill."""

    def __init__(self, nseq=16, L=2048, Hq=32, Hkv=8, D=128, seed=0, device="cuda", dtype="bf16"):
        import torch

        tdt = torch.bfloat16 if dtype == "bf16" else torch.float16

        self.nseq, self.L, self.Hq, self.Hkv, self.D = nseq, L, Hq, Hkv, D
        n = nseq * L
        self.n = n
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        self.q = torch.randn((n, Hq, D), generator=g, device=device, dtype=tdt)
        self.k = torch.randn((n, Hkv, D), generator=g, device=device, dtype=tdt)
        self.v = torch.randn((n, Hkv, D), generator=g, device=device, dtype=tdt)
        ip = (np.arange(nseq + 1) * L).astype(np.int32)
        self.indptr = torch.from_numpy(ip).to(device)
        self.qpos = torch.from_numpy(np.tile(np.arange(L, dtype=np.int32), nseq)).to(device)
        self.kofs = torch.zeros(nseq, dtype=torch.int32, device=device)
        self.o = torch.empty((n, Hq, D), device=device, dtype=tdt)
        self.lse = torch.empty((n, Hq), device=device, dtype=torch.float32)
        self.sm_scale = D ** -0.5

    def flops(self):
        return causal_prefill_flops_of(self.nseq, self.L, self.Hq, self.D)

    def bytes(self):
        return (2 * self.n * self.Hq + 2 * self.n * self.Hkv) * self.D * 2 + 4 * self.n * self.Hq

    def run(self, capi):
        capi.attention_prefill_ragged(self.q, self.indptr, self.k, self.v, self.indptr, self.qpos, self.kofs, self.o,
                                      self.lse, 1, 0, 1.0, 5e5, self.sm_scale)


class AppendWorkload:
    """C3's KV append: the 16 x 2048 new tokens go into freshly allocated (permuted) pages."""

    def __init__(self, nseq=16, L=2048, Hq=32, Hkv=8, D=128, page=16, seed=0, device="cuda"):
        import torch

        self.n, self.Hq, self.Hkv, self.D = nseq * L, Hq, Hkv, D
        n = self.n
        rng = np.random.default_rng(seed)
        P = n // page + 1
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        self.qkv = torch.randn((n, Hq + 2 * Hkv, D), generator=g, device=device, dtype=torch.bfloat16)
        self.q = torch.empty((n, Hq, D), device=device, dtype=torch.bfloat16)
        self.k = torch.randn((n, Hkv, D), generator=g, device=device, dtype=torch.bfloat16)
        self.v = torch.randn((n, Hkv, D), generator=g, device=device, dtype=torch.bfloat16)
        self.pages = torch.zeros((P, 2, Hkv, page, D), device=device, dtype=torch.bfloat16)
        perm = rng.permutation(P).astype(np.int32)[: n // page]
        slots = (perm[:, None] * page + np.arange(page, dtype=np.int32)[None, :]).reshape(-1).astype(np.int32)
        self.slots = torch.from_numpy(slots).to(device)
        self.pos = torch.from_numpy(np.tile(np.arange(L, dtype=np.int32), nseq)).to(device)

    def append_bytes(self):
        return append_bytes_of(self.n, self.Hkv, self.D)

    def rotary_append_bytes(self):
        # qkv read once; q, k, v written; k, v written into the pages; the two position arrays
        e = 2
        return (2 * self.n * (self.Hq + 2 * self.Hkv) * self.D * e + self.n * self.Hkv * self.D * 2 * e + 8 * self.n)

    def run_append(self, capi):
        capi.transpose_append(self.pages, self.k, self.v, self.slots)

    def run_rotary_append(self, capi):
        capi.split_rotary_append(self.qkv, self.pos, self.slots, self.q, self.k, self.v, self.pages, 1, 1.0, 5e5)


def dfs_tree_mask(parents):
    """(dfs order, subtree end) per node, as the reference's ConstructTokenTreeMask emits (paged_kv_cache.cc:1900-1918)"""
    n = len(parents)
    children = [[] for _ in range(n)]
    roots = []
    for i, p in enumerate(parents):
        (roots if p < 0 else children[p]).append(i)
    order, end, cnt = [0] * n, [0] * n, [0]

    def visit(u):
        order[u] = cnt[0]
        cnt[0] += 1
        for c in children[u]:
            visit(c)
        end[u] = cnt[0]

    for r in roots:
        visit(r)
    return np.array([[order[i], end[i]] for i in range(n)], np.int32)


class TreeWorkload:
    """C5 (i): B sequences with L committed tokens each; a `nodes`-node token tree per sequence is verified in one round.
    First-round callback sequence of the cache (kv_cache_host.cc / paged_kv_cache.cc MHASelfAttnInternal +
    MHACrossAttnInternal): tree-masked ragged self-attention over the new nodes, mask-free paged prefill over the
    committed context, f_merge_inplace."""

    def __init__(self, B=32, L=32768, nodes=64, Hq=32, Hkv=8, D=128, page=16, seed=0, device="cuda", dtype="bf16"):
        import torch

        tdt = torch.bfloat16 if dtype == "bf16" else torch.float16
        self.B, self.L, self.nodes, self.Hq, self.Hkv, self.D = B, L, nodes, Hq, Hkv, D
        rng = np.random.default_rng(seed)
        n = B * nodes
        self.n = n
        ppseq = L // page
        self.nnz = B * ppseq
        P = self.nnz + 1
        g = torch.Generator(device=device)
        g.manual_seed(seed)
        self.pages = torch.randn((P, 2, Hkv, page, D), generator=g, device=device, dtype=tdt)
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, np.int32)).to(device)  # noqa: E731
        self.page_values = i32(rng.permutation(P)[: self.nnz])
        self.page_indptr = i32(np.arange(B + 1) * ppseq)
        self.length_info = i32(np.full(B, page))
        self.kofs = i32(np.zeros(B))
        # even sequences: the complete binary tree of BASELINE.md C5; odd ones: random parents (seed 0)
        trees = [[(i - 1) // 2 if i else -1 for i in range(nodes)] if b % 2 == 0 else
                 [-1] + [int(rng.integers(0, i)) for i in range(1, nodes)] for b in range(B)]
        depth = []
        self.pairs = 0
        for t in trees:
            d = []
            for i, p in enumerate(t):
                d.append(0 if p < 0 else d[p] + 1)
            depth.append(d)
            self.pairs += sum(x + 1 for x in d)
        self.mask = i32(np.concatenate([dfs_tree_mask(t) for t in trees]))
        self.indptr = i32(np.arange(B + 1) * nodes)
        self.qpos = i32(np.concatenate([L + np.array(d) for d in depth]))
        self.q = torch.randn((n, Hq, D), generator=g, device=device, dtype=tdt)
        self.k = torch.randn((n, Hkv, D), generator=g, device=device, dtype=tdt)
        self.v = torch.randn((n, Hkv, D), generator=g, device=device, dtype=tdt)
        self.o = torch.empty((n, Hq, D), device=device, dtype=tdt)
        self.lse = torch.empty((n, Hq), device=device, dtype=torch.float32)
        self.o2, self.lse2 = torch.empty_like(self.o), torch.empty_like(self.lse)
        self.sm_scale = D ** -0.5

    def flops(self):
        return 4 * self.D * self.Hq * (self.n * self.L + self.pairs)

    def bytes(self):
        e = 2
        return self.B * self.L * self.Hkv * self.D * 2 * e + 4 * self.n * self.Hq * self.D * e

    def run(self, capi):
        capi.attention_prefill_tree_ragged(self.q, self.indptr, self.k, self.v, self.indptr, self.qpos, self.indptr,
                                           self.mask, self.o, self.lse, 0, 1.0, 5e5, self.sm_scale)
        capi.attention_prefill_paged(self.q, self.indptr, self.pages, self.page_indptr, self.page_values,
                                     self.length_info, self.kofs, self.qpos, self.o2, self.lse2, 0, 0, 1.0, 5e5,
                                     self.sm_scale)
        capi.merge_state_inplace(self.o, self.lse, self.o2, self.lse2)


# ---------------------------------------------------------------------------------------------------
# clocks sampling (NVML) during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index=0, period_s=0.005):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.period, self._stop, self._thr, self.h = period_s, threading.Event(), None, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        if self.h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": int(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# timing
# ---------------------------------------------------------------------------------------------------
def time_windows(step, K, R, dev, dist=None, min_window_ms=0.0):
    """R windows of exactly K steps each: CUDA events on the launching stream around the window, a device synchronize
    (and, multi-rank, a barrier) on both sides, the max over ranks per window.  Returns (median ms per step, the list of
    per-window ms per step, clocks summary over all windows).  Nothing but the steps' own launches is on the stream
    inside a window (an event between the kernels would break their programmatic dependent launch)."""
    import torch

    beg = [torch.cuda.Event(enable_timing=True) for _ in range(R)]
    end = [torch.cuda.Event(enable_timing=True) for _ in range(R)]
    with ClockSampler(dev.index or 0) as clk:
        for r in range(R):
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
                torch.cuda.synchronize()
            beg[r].record()
            for _ in range(K):
                step()
            end[r].record()
            torch.cuda.synchronize()
    ms = [beg[r].elapsed_time(end[r]) for r in range(R)]
    if dist is not None:
        t = torch.tensor(ms, device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = [float(x) for x in t.tolist()]
    per_step = sorted(m / K for m in ms)
    return per_step[len(per_step) // 2], [round(m / K, 5) for m in ms], clk.summary()


def cupti_kernels(fn, iters=3):
    """Per-kernel device durations of `iters` calls of fn from CUPTI (torch.profiler / kineto activity records): name ->
    {count per call, avg us}.  Informational (which kernels a step launches and their share); with programmatic
    dependent launch a dependent kernel's record includes the time it waits for its primary, so these are never used as a
    roofline denominator.  Returns None if the profiler is unavailable."""
    try:
        import torch
        from torch.autograd import DeviceType
        from torch.profiler import ProfilerActivity, profile

        fn()
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(iters):
                fn()
            torch.cuda.synchronize()
        out = {}
        for e in prof.key_averages():
            if getattr(e, "device_type", None) != DeviceType.CUDA:
                continue
            name = e.key
            if name.startswith("Memcpy") or name.startswith("Memset"):
                continue
            tot = getattr(e, "device_time_total", None)
            if tot is None:
                tot = getattr(e, "cuda_time_total", 0.0)
            short = name.split("<")[0].split("(")[0].split("::")[-1]
            d = out.setdefault(short, {"per_step": 0.0, "avg_us": 0.0, "_tot": 0.0, "_cnt": 0})
            d["_tot"] += float(tot)
            d["_cnt"] += int(e.count)
        for d in out.values():
            d["per_step"] = round(d["_cnt"] / iters, 2)
            d["avg_us"] = round(d["_tot"] / max(d["_cnt"], 1), 2)
            del d["_tot"], d["_cnt"]
        return out or None
    except Exception as e:  # pragma: no cover
        print(f"[bench] CUPTI kernel list unavailable: {e!r}", file=sys.stderr)
        return None


def dist_setup(n_gpus):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def cpu_decode_baseline(L=4096, Hq=32, Hkv=8, D=128, B=1, repeats=50):
    """Times the CPU decode kernel on a bounded slice (B sequences of the C2 shape, fp16 like the
    reference's CPU tests).  Returns dict(value GB/s, unit, cores, kind, sample)."""
    from oracle import cpu_ref

    return cpu_ref.time_decode(B=B, L=L, Hq=Hq, Hkv=Hkv, D=D, repeats=repeats)


def _free():
    import gc

    import torch

    gc.collect()
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------------------------------
# sub-records
# ---------------------------------------------------------------------------------------------------
def sub_prefill_c3(args, capi, dev, peaks, peak_src, dtype="bf16", seed=0):
    import torch

    w = PrefillWorkload(seed=seed, device=dev, dtype=dtype)
    for _ in range(5):
        w.run(capi)
    torch.cuda.synchronize()
    # 15 windows of 20 launches, ~0.2 s of tensor work back to back: a B200 starts this kernel at its boost clock and
    # settles 10-15 % lower under its power cap within ~100 ms (the windows show it; MEASURED_PEAKS' burst / sustained
    # cuBLAS figures differ by the same ratio).  `value` is the median window; `burst` (fastest window) and `sustained`
    # (median of the last five) are each compared with the peak of their own kind.
    K, R = 20, 15
    n0 = capi.launch_count()
    ms, windows, clocks = time_windows(lambda: w.run(capi), K, R, dev)
    launches = (capi.launch_count() - n0) // R
    kern = None if args.no_cupti else cupti_kernels(lambda: w.run(capi))
    tf = w.flops() / (ms * 1e-3) / 1e12
    tf_burst = w.flops() / (min(windows) * 1e-3) / 1e12
    tf_sus = w.flops() / (sorted(windows[-5:])[2] * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"]))
    sus = peaks.get("bf16_tflops_sustained")
    traffic, tsrc = ncu_traffic("prefill_c3")
    del w
    _free()
    return {"metric": "prefill_tflops", "value": round(tf, 2), "unit": "TFLOP/s", "ms_per_step": round(ms, 4), "steps": K,
            "windows": R, "windows_ms": windows, "dtype": dtype,
            "config": {"workload": "C3 ragged causal prefill 16x2048, 32q/8kv heads, D128 (f_attention_prefill_ragged)",
                       "l2": "q/k/v/o 0.67 GB > 126 MB L2"},
            "roofline": {"bound": "tensor", "kernel": "prefill_tc05_kernel", "achieved": round(tf, 2), "peak": peak,
                         "unit": "TFLOP/s", "frac": round(tf / peak, 4), "peak_source": f"of {peak_src} (burst)",
                         "burst": round(tf_burst, 2), "burst_frac_of_burst_peak": round(tf_burst / peak, 4),
                         "sustained": round(tf_sus, 2), "peak_sustained": sus,
                         "sustained_frac_of_sustained_peak": round(tf_sus / float(sus), 4) if sus else None,
                         "frac_of_spec_2250": round(tf / 2250.0, 4), "algorithmic_flops": w_flops_c3(),
                         "traffic": traffic, "traffic_source": tsrc},
            "gpu_launches": int(launches), "kernels": kern, "clocks": clocks}


def w_flops_c3():
    return causal_prefill_flops_of(16, 2048, 32, 128)


def sub_append_c3(args, capi, dev, peaks, peak_src):
    import torch

    w = AppendWorkload(device=dev)
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))
    out = {"metric": "append_hbm_gbps", "unit": "GB/s",
           "config": {"workload": "C3 KV append: 32768 new tokens (16 x 2048) into permuted pages, 8 kv heads, D128, bf16",
                      "l2": "k/v in + pages out 268 MB (805 MB with the rotary's qkv / q / k / v) > 126 MB L2"}}
    K, R = 20, 7
    for name, fn, nbytes, kname in (("transpose_append", w.run_append, w.append_bytes(), "transpose_append_kernel"),
                                    ("split_rotary_append", w.run_rotary_append, w.rotary_append_bytes(),
                                     "split_rotary_kernel<APPEND>")):
        for _ in range(5):
            fn(capi)
        torch.cuda.synchronize()
        n0 = capi.launch_count()
        ms, windows, clocks = time_windows(lambda: fn(capi), K, R, dev)
        gbs = nbytes / (ms * 1e-3) / 1e9
        traffic, tsrc = ncu_traffic("append_c3_" + name)
        out[name] = {"value": round(gbs, 1), "ms_per_step": round(ms, 5), "steps": K, "windows": R, "windows_ms": windows,
                     "roofline": {"bound": "hbm", "kernel": kname, "achieved": round(gbs, 1), "peak": hbm_peak,
                                  "unit": "GB/s", "frac": round(gbs / hbm_peak, 4), "peak_source": f"of {peak_src}",
                                  "frac_of_spec_8000": round(gbs / 8000.0, 4), "algorithmic_bytes": int(nbytes),
                                  "traffic": traffic, "traffic_source": tsrc},
                     "gpu_launches": int((capi.launch_count() - n0) // R), "clocks": clocks}
    out["value"] = out["transpose_append"]["value"]
    del w
    _free()
    return out


def sub_c5_tree_prefill(args, capi, dev, peaks, peak_src, dtype="bf16"):
    import torch

    w = TreeWorkload(device=dev, dtype=dtype)
    for _ in range(3):
        w.run(capi)
    torch.cuda.synchronize()
    K, R = 5, 5
    n0 = capi.launch_count()
    ms, windows, clocks = time_windows(lambda: w.run(capi), K, R, dev)
    launches = (capi.launch_count() - n0) // R
    kern = None if args.no_cupti else cupti_kernels(lambda: w.run(capi), iters=2)
    tf = w.flops() / (ms * 1e-3) / 1e12
    gbs = w.bytes() / (ms * 1e-3) / 1e9
    peak = float(peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"]))
    sus = peaks.get("bf16_tflops_sustained")
    traffic, tsrc = ncu_traffic("c5_tree_prefill")
    out = {"metric": "prefill_tflops", "value": round(tf, 2), "unit": "TFLOP/s", "ms_per_step": round(ms, 4), "steps": K,
           "windows": R, "windows_ms": windows, "dtype": dtype,
           "config": {"workload": "C5 tree prefill: batch 32 x 64-node token trees (complete binary / random parents) over "
                                  "32768 cached tokens each, 32q/8kv, D128, page16; step = tree-masked ragged self part + "
                                  "mask-free paged part over the cache + merge (first tree round of the cache)",
                      "l2": "KV working set 4.3 GB > 126 MB L2"},
           "roofline": {"bound": "tensor", "kernel": "prefill_tc05_kernel<PAGED>", "achieved": round(tf, 2), "peak": peak,
                        "unit": "TFLOP/s", "frac": round(tf / peak, 4), "peak_source": f"of {peak_src} (burst)",
                        "frac_of_sustained": round(tf / float(sus), 4) if sus else None,
                        "algorithmic_flops": int(w.flops()), "algorithmic_bytes": int(w.bytes()),
                        "hbm_gbs": round(gbs, 1), "traffic": traffic, "traffic_source": tsrc},
           "gpu_launches": int(launches), "kernels": kern, "clocks": clocks}
    del w
    _free()
    return out


def sub_c5_decode(args, capi, dev, peaks, peak_src):
    import torch

    w = DecodeWorkload(B=32, L=32768, device=dev)
    for _ in range(5):
        w.run_step_fused(capi)
    torch.cuda.synchronize()
    K, R = 20, 7
    n0 = capi.launch_count()
    ms, windows, clocks = time_windows(lambda: w.run_step_fused(capi), K, R, dev)
    launches = (capi.launch_count() - n0) // R
    gbs = w.step_bytes() / (ms * 1e-3) / 1e9
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))
    traffic, tsrc = ncu_traffic("c5_decode_32k")
    out = {"metric": "decode_attn_hbm_gbps", "value": round(gbs, 1), "unit": "GB/s", "ms_per_step": round(ms, 5), "steps": K,
           "windows": R, "windows_ms": windows, "dtype": "bf16", "tok_s_layer": round(w.B / (ms * 1e-3), 1),
           "config": {"workload": "C5 split-KV decode: batch 32 x 32768 ctx, 32q/8kv heads, D128, page16, bf16; step = "
                                  "split_rotary+append+decode of one layer (one fused launch + the split-KV merge)",
                      "l2": "KV working set 4.3 GB > 126 MB L2"},
           "roofline": {"bound": "hbm", "kernel": "decode_kernel<FUSED qkv>(+decode_merge_kernel)", "achieved": round(gbs, 1),
                        "peak": hbm_peak, "unit": "GB/s", "frac": round(gbs / hbm_peak, 4), "peak_source": f"of {peak_src}",
                        "frac_of_spec_8000": round(gbs / 8000.0, 4), "algorithmic_bytes": int(w.step_bytes()),
                        "traffic": traffic, "traffic_source": tsrc},
           "gpu_launches": int(launches), "clocks": clocks}
    del w
    _free()
    return out


def sub_ref_gpu(args, our):
    """The reference's own GPU TIR kernels (oracle/_ref/ref_gpu_kernels_bfloat16_*.so inside the reference's runtime, a
    subprocess) on the C2 decode and C3 ragged-prefill inputs: the kernel to beat.  A baseline leg like cpu_baseline:
    nothing of the product path runs here."""
    from oracle import ref_gpu

    if not ref_gpu.available("bfloat16"):
        return {"unavailable": "oracle/_ref/tvm_cuda or the bf16 kernel module is not packed (oracle/ref_harness/pack_ref_cuda.sh)"}
    B, L, Hq, Hkv, D, page = 64, 4096, 32, 8, 128, 16
    rng = np.random.default_rng(0)
    ppseq = L // page
    nnz = B * ppseq
    P = nnz + 1
    n3, L3 = 16 * 2048, 2048
    ip3 = (np.arange(17) * L3).astype(np.int32)
    arrays = {"iv": rng.permutation(P).astype(np.int32)[:nnz], "ip": (np.arange(B + 1) * ppseq).astype(np.int32),
              "li": np.full(B, page, np.int32), "kro": np.zeros(B, np.int32), "qpos": np.full(B, L - 1, np.int32),
              "qi3": ip3, "ki3": ip3.copy(), "qpos3": np.tile(np.arange(L3, dtype=np.int32), 16),
              "kro3": np.zeros(16, np.int32)}
    tensors = {k: {"dtype": "int32", "init": "npz"} for k in arrays}
    bf = "bfloat16"
    tensors.update({
        "pages": {"dtype": bf, "shape": [P, 2, Hkv, page, D], "init": "randn"},
        "q": {"dtype": bf, "shape": [B, Hq, D], "init": "randn"},
        "o": {"dtype": bf, "shape": [B, Hq, D], "init": "zeros"},
        "lse": {"dtype": "float32", "shape": [B, Hq], "init": "zeros"},
        "q3": {"dtype": bf, "shape": [n3, Hq, D], "init": "randn"},
        "k3": {"dtype": bf, "shape": [n3, Hkv, D], "init": "randn"},
        "v3": {"dtype": bf, "shape": [n3, Hkv, D], "init": "randn"},
        "o3": {"dtype": bf, "shape": [n3, Hq, D], "init": "zeros"},
        "lse3": {"dtype": "float32", "shape": [n3, Hq], "init": "zeros"},
    })
    sm = D ** -0.5
    calls = [{"fn": "batch_decode_paged_kv", "args": ["q", "pages", "ip", "iv", "li", "kro", "qpos", "o", "lse", 0, 1.0, 5e5, sm]},
             {"fn": "batch_prefill_ragged_kv", "args": ["q3", "qi3", "k3", "v3", "ki3", "qpos3", "kro3", "o3", "lse3", 1, 0, 1.0, 5e5, sm]}]
    spec = {"module": ref_gpu.kernel_module("bfloat16").name, "tensors": tensors, "calls": calls, "fetch": [],
            "time": [{"call": 0, "warmup": 3, "iters": 20}, {"call": 1, "warmup": 2, "iters": 5}]}
    res, _ = ref_gpu.run_kernels(spec, arrays, timeout=600)
    if not res.get("ok"):
        return {"error": str(res.get("error"))[:300]}
    t = res["timings_ms"]
    dec_bytes = B * L * Hkv * D * 2 * 2 + 2 * B * Hq * D * 2 + 4 * B * Hq + 4 * (nnz + 4 * B + 1)
    d_ms, p_ms = t["batch_decode_paged_kv"]["median"], t["batch_prefill_ragged_kv"]["median"]
    out = {"what": "the reference's own GPU TIR kernels built for sm_100a (bf16), CUDA events inside the reference's runtime",
           "decode_c2": {"kernel": "batch_decode_paged_kv (_decode_kernels.py:181-411)", "ms": round(d_ms, 4),
                         "gbs": round(dec_bytes / (d_ms * 1e-3) / 1e9, 1), "note": "f_attention_decode alone (no rotary / append)"},
           "prefill_c3": {"kernel": "batch_prefill_ragged_kv (_prefill_kernels.py:795-923)", "ms": round(p_ms, 3),
                          "tflops": round(w_flops_c3() / (p_ms * 1e-3) / 1e12, 2)}}
    if our.get("decode_ms"):
        out["decode_c2"]["ours_ms"] = round(our["decode_ms"], 4)
        out["decode_c2"]["speedup"] = round(d_ms / our["decode_ms"], 2)
    if our.get("prefill_ms"):
        out["prefill_c3"]["ours_ms"] = round(our["prefill_ms"], 4)
        out["prefill_c3"]["speedup"] = round(p_ms / our["prefill_ms"], 2)
    return out


def sub_c4(args, capi, rank, world, dev, peaks, peak_src, dist):
    """C4: Llama-3-70B GQA decode (64 q / 8 kv heads), batch 256 at 8K context, KV-head groups sharded across the ranks
    (strong scaling: total work fixed); the per-head outputs are re-assembled on every rank inside the timed step."""
    import torch

    from tvm_b200 import sharding

    Hq, Hkv, B, L = 64, 8, 256, 8192
    if Hkv % world:
        return {"skipped": f"8 kv heads do not split over {world} ranks"}
    q0, q1, k0, k1 = sharding.head_shard(Hq, Hkv, world, rank)
    w = DecodeWorkload(B=B, L=L, Hq=q1 - q0, Hkv=k1 - k0, seed=0, device=dev)  # same page table on every rank
    gather, gather_kind = None, "none"
    if world > 1:
        gather_kind = args.gather
        if gather_kind in ("p2p", "p2p-unfused"):
            try:
                gather = sharding.PeerHeadGather(B, Hq, w.D, torch.bfloat16, dev)
            except Exception as e:  # symmetric memory unavailable on this box: the NCCL all-gather still works
                if rank == 0:
                    print(f"[bench] peer gather unavailable ({e!r}); using the NCCL all-gather", file=sys.stderr)
                gather_kind = "nccl"
        if gather is not None:
            # parity of the fused path against decode + NCCL all-gather on the same inputs, once, before timing
            got = gather.decode_fused_qkv(capi, w.qkv, w.q_rope_position, w.append_position, w.pages, w.page_indptr,
                                          w.page_values, w.length_info, w.k_rope_pos_offset, w.o, w.lse, 1, w.rope_scale,
                                          w.rope_theta, w.sm_scale).clone()
            w.run_rotary_append(capi)
            w.run_decode(capi)
            want = sharding.all_gather_heads(w.o)
            torch.cuda.synchronize()
            assert torch.allclose(got.float(), want.float(), atol=2e-3, rtol=1e-2), "peer-gathered heads differ from decode + NCCL all-gather"

    def make_step(kind):
        def step():
            if gather is not None and kind == "p2p-unfused":  # rotary+append launch, then decode + head gather
                w.run_rotary_append(capi)
                return w.run_decode_gather(capi, gather)
            if gather is not None and kind == "p2p":  # rotary + append + decode + head gather: one fused launch (+ merge / wait)
                return gather.decode_fused_qkv(capi, w.qkv, w.q_rope_position, w.append_position, w.pages, w.page_indptr,
                                               w.page_values, w.length_info, w.k_rope_pos_offset, w.o, w.lse, 1,
                                               w.rope_scale, w.rope_theta, w.sm_scale)
            w.run_step_fused(capi)
            if world > 1 and kind != "local":
                return sharding.all_gather_heads(w.o)
            return w.o
        return step

    K, R = max(args.steps, 20), 7
    total_bytes = world * w.step_bytes()
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))
    res = {}
    kinds = [gather_kind] + (["nccl"] if world > 1 and gather_kind != "nccl" else []) + (["local"] if world > 1 else [])
    for kind in kinds:
        step = make_step(kind)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        n0 = capi.launch_count()
        ms, windows, clocks = time_windows(step, K, R, dev, dist if world > 1 else None)
        res[kind] = {"ms_per_step": round(ms, 5), "value": round(total_bytes / (ms * 1e-3) / 1e9, 1), "windows_ms": windows,
                     "per_rank_hbm_frac": round(total_bytes / world / (ms * 1e-3) / 1e9 / hbm_peak, 4),
                     "gpu_launches": int((capi.launch_count() - n0) // R), "clocks": clocks}
    main = res[gather_kind]
    out = {"metric": "decode_attn_hbm_gbps", "value": main["value"], "unit": "GB/s", "n_gpus": world, "steps": K, "windows": R,
           "ms_per_step": main["ms_per_step"], "scaling": "strong", "dtype": "bf16",
           "config": {"workload": "C4 Llama-3-70B GQA decode: batch 256 x 8192 ctx, 64q/8kv heads sharded by KV-head group, "
                                  "D128, page16, bf16; step = split_rotary+append+decode+re-assembly of the per-head O on every rank",
                      "head_gather": {"p2p": "in-kernel NVLink peer stores + flags, fused qkv step (tvmb200_attention_decode_fused_qkv_gather)",
                                      "p2p-unfused": "in-kernel NVLink peer stores + flags (tvmb200_attention_decode_gather)",
                                      "nccl": "ncclAllGather behind the kernel", "none": "single GPU"}[gather_kind],
                      "global_batch": B, "seq_len": L, "parallelism": f"tp{world} (KV-head groups)",
                      "l2": "KV working set >= 1 GiB/GPU > 126 MB L2"},
           "tok_s_layer": round(B / (main["ms_per_step"] * 1e-3), 1),
           "roofline": {"bound": "hbm", "achieved": round(main["value"] / world, 1), "peak": hbm_peak, "unit": "GB/s",
                        "frac": main["per_rank_hbm_frac"], "traffic": None, "peak_source": f"of {peak_src}",
                        "algorithmic_bytes": int(total_bytes // world),
                        "note": "per-rank share of the whole step incl. the head re-assembly; gather payload "
                                f"{B * Hq * w.D * 2 * (world - 1) // max(world, 1)} B received per rank per step"},
           "gather_kinds": res, "gpu_launches": main["gpu_launches"], "clocks": main["clocks"]}
    del w, gather
    _free()
    return out


# ---------------------------------------------------------------------------------------------------
# headline
# ---------------------------------------------------------------------------------------------------
def run_own(args):
    import torch

    from tvm_b200 import capi

    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    capi.lib()
    peaks, peak_src = measured_peaks()
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: F811

    def finish(rec):
        if rank == 0:
            rec.setdefault("n_gpus", world)
            rec.setdefault("warmup", max(args.warmup, 3))
            rec.setdefault("higher_is_better", True)
            rec.setdefault("scaling", "weak")
            rec.setdefault("vs_baseline", None)
            rec.setdefault("data", "synthetic")
            print(json.dumps(rec), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    single = {"prefill": lambda: sub_prefill_c3(args, capi, dev, peaks, peak_src, dtype=args.dtype, seed=rank),
              "append": lambda: sub_append_c3(args, capi, dev, peaks, peak_src),
              "c5": lambda: sub_c5_tree_prefill(args, capi, dev, peaks, peak_src, dtype=args.dtype),
              "c5decode": lambda: sub_c5_decode(args, capi, dev, peaks, peak_src),
              "c4": lambda: sub_c4(args, capi, rank, world, dev, peaks, peak_src, dist)}
    if args.workload in single:
        return finish(single[args.workload]())

    w = DecodeWorkload(seed=rank, device=dev)

    def step():
        w.run_step_fused(capi)

    W = max(args.warmup, 3)
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    K, R = args.steps, args.windows
    n0 = capi.launch_count()
    ms_per_step, windows, clocks = time_windows(step, K, R, dev, dist)
    launches = (capi.launch_count() - n0) // R
    kern = None if (args.no_cupti or rank != 0) else cupti_kernels(step, iters=5)
    # f_attention_decode alone (q already rotated, nothing appended): the like-for-like number next to ref_gpu's
    dec_ms = None
    if world == 1 and not args.no_sub:
        w.run_rotary_append(capi)
        for _ in range(3):
            w.run_decode(capi)
        dec_ms, _, _ = time_windows(lambda: w.run_decode(capi), 20, 5, dev)
    value = world * w.step_bytes() / (ms_per_step * 1e-3) / 1e9
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))
    # roofline of the dominant launch pair: the step IS that pair (fused decode launch + its split-KV merge, nothing else
    # on the stream), so its average duration is the window time / K -- events around each pair would break the
    # programmatic dependent launch between consecutive steps and over-state it
    per_rank_gbs = w.step_bytes() / (ms_per_step * 1e-3) / 1e9
    traffic, tsrc = ncu_traffic("decode_c2")
    out = {
        "metric": "decode_attn_hbm_gbps", "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "windows": R, "windows_ms": windows,
        "config": {"workload": "C2 Llama-3-8B decode: batch 64/GPU x 4096 ctx, 32q/8kv heads, D128, page16, bf16 "
                               "paged KV; step = split_rotary+append+decode of one layer (one fused launch + the split-KV merge)",
                   "global_batch": w.B * world, "seq_len": w.L, "parallelism": f"batch-split x{world} (no data-path collective)",
                   "timing": f"median of {R} windows of exactly {K} steps (max over ranks per window)",
                   "l2": "KV working set 1 GiB/GPU > 126 MB L2 (no flush needed)"},
        "tok_s_layer": round(world * w.B / (ms_per_step * 1e-3), 1),
        "roofline": {"bound": "hbm", "kernel": "decode_kernel<FUSED qkv>(+decode_merge_kernel)", "achieved": round(per_rank_gbs, 1),
                     "peak": hbm_peak, "unit": "GB/s", "frac": round(per_rank_gbs / hbm_peak, 4), "traffic": traffic,
                     "traffic_source": tsrc, "peak_source": f"of {peak_src}", "algorithmic_bytes": w.step_bytes(),
                     "kernel_ms": round(ms_per_step, 5),
                     "kernel_ms_source": "timed region / launch pairs (CUDA events over the windows)",
                     "frac_of_spec_8000": round(per_rank_gbs / 8000.0, 4)},
        "gpu_launches": int(launches), "kernels": kern,
        "clocks": clocks,
    }
    try:
        e2e = {"value": None, "skipped": "--no-e2e"} if args.no_e2e else run_e2e(args, w, world, dist)
    except Exception as e:  # pragma: no cover
        e2e = {"value": None, "error": repr(e)[:300]}
    out["e2e"] = e2e
    del w
    _free()
    subs = {}
    if not args.no_sub:
        def guarded(name, fn):
            t0 = time.time()
            try:
                r = fn()
            except Exception as e:  # a failing sub-record must not take the headline down
                r = {"error": repr(e)[:300]}
                _free()
            if isinstance(r, dict):
                r["wall_s"] = round(time.time() - t0, 1)
            subs[name] = r

        if world == 1:
            guarded("prefill_c3", lambda: sub_prefill_c3(args, capi, dev, peaks, peak_src))
            guarded("append_c3", lambda: sub_append_c3(args, capi, dev, peaks, peak_src))
            guarded("c5_tree_prefill", lambda: sub_c5_tree_prefill(args, capi, dev, peaks, peak_src))
            guarded("c5_decode_32k", lambda: sub_c5_decode(args, capi, dev, peaks, peak_src))
            if not args.no_cpu:
                ours = {"decode_ms": dec_ms, "prefill_ms": (subs.get("prefill_c3") or {}).get("ms_per_step")}
                guarded("ref_gpu", lambda: sub_ref_gpu(args, ours))
        else:
            guarded("c4_head_sharded", lambda: sub_c4(args, capi, rank, world, dev, peaks, peak_src, dist))
    out["sub"] = subs
    if rank == 0:
        if world == 1 and not args.no_cpu:
            try:
                out["cpu_baseline"] = cpu_decode_baseline()
            except Exception as e:  # pragma: no cover
                out["cpu_baseline"] = {"value": None, "error": repr(e)[:200]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(args, w, world, dist=None):
    """End to end through the repo's public API -- the C++ host cache (tvm_b200.kv_cache.PagedKVCache, the drop-in for
    the reference's PagedAttentionKVCacheObj) -- with HOST inputs: every step does begin_forward (host page-table
    bookkeeping), copies the step's fused qkv from pinned host memory, runs attention_with_fused_qkv (one merged H2D
    copy of the aux arrays + split_rotary + append + decode on the sm_100a kernels), copies O back to pinned host
    memory.  The context grows by one token per step (4096 .. 4096+K) and the byte count follows it."""
    import torch

    from tvm_b200.kv_cache import PagedKVCache

    dev = w.qkv.device
    B, L, Hq, Hkv, D = w.B, w.L, w.Hq, w.Hkv, w.D
    chunk = 8192
    K = max(100, args.steps)
    cache = PagedKVCache(reserved_num_seqs=B, total_token_capacity=B * (L + K + 32), prefill_chunk_size=chunk, num_layers=1,
                         num_qo_heads=Hq, num_kv_heads=Hkv, head_dim=D, rope_mode=1, rotary_theta=w.rope_theta,
                         dtype="bfloat16", device=dev.index or 0)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    fill = torch.randn((chunk, Hq + 2 * Hkv, D), generator=g, device=dev, dtype=torch.bfloat16)
    fill_o = torch.empty((chunk, Hq, D), device=dev, dtype=torch.bfloat16)
    for sid in range(B):  # build the 4095-token context of every sequence through the cache itself (chunked prefill)
        cache.add_sequence(sid)
        left = L - 1
        while left > 0:
            n = min(left, chunk)
            cache.begin_forward([sid], [n])
            cache.attention_with_fused_qkv(0, w.sm_scale, fill[:n], fill_o[:n])
            cache.end_forward()
            left -= n
    torch.cuda.synchronize()
    seq_ids, ones = list(range(B)), [1] * B
    # Host buffers in, host buffers out, every step -- pipelined the way a serving loop would: the pinned-host -> device
    # copy of step i+1's qkv and the device -> pinned-host copy of step i's O run on a copy stream while step i / i+1
    # compute; two device buffers each, events for the hand-offs.  All copies are inside the timed region.
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    d_qkv = [torch.empty_like(w.qkv) for _ in range(2)]
    d_o = [torch.empty_like(w.o) for _ in range(2)]
    h_out = torch.empty(w.o.shape, dtype=torch.bfloat16).pin_memory()
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]   # compute of the step that last used buffer pair i is done

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[i & 1])
            d_qkv[i & 1].copy_(w.h_qkv, non_blocking=True)
            ev_in[i & 1].record(copy_stream)

    def run_steps(n, first):
        upload(first)
        for i in range(first, first + n):
            if i + 1 < first + n:
                upload(i + 1)
            cache.begin_forward(seq_ids, ones)
            main.wait_event(ev_in[i & 1])
            cache.attention_with_fused_qkv(0, w.sm_scale, d_qkv[i & 1], d_o[i & 1])
            ev_out[i & 1].record(main)
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_out[i & 1])
                h_out.copy_(d_o[i & 1], non_blocking=True)
                ev_free[i & 1].record(copy_stream)
            cache.end_forward()

    for e in ev_free:
        e.record(main)
    run_steps(3, 0)  # context is now L + 2 after warm-up; every timed step appends one more token per sequence
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(K, 3)
    main.wait_stream(copy_stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    aux_ints = B + 3 * (B + 1) + w.nnz + 2 * B + B + 2 * B  # q_rope, indptrs, page ids, len/rope arrays, append map
    del cache
    # algorithmic bytes with the real context lengths: step s (0-based, after 3 warm-up steps) reads L + 3 + s tokens
    extra_kv = (3 + (K - 1) / 2.0) * B * Hkv * D * 2 * 2
    return {"value": round(world * (w.step_bytes() + extra_kv) / (ms * 1e-3) / 1e9, 1), "unit": "GB/s",
            "h2d_bytes_per_step": int(w.h_qkv.numel() * 2 + aux_ints * 4), "d2h_bytes_per_step": int(h_out.numel() * 2),
            "ms_per_step": round(ms, 5), "steps": K,
            "api": "tvm_b200.kv_cache.PagedKVCache.begin_forward/attention_with_fused_qkv (C ABI tvmb200_cache_*); "
                   "H2D of the next step and D2H of the previous one overlap the compute on a copy stream"}


def run_reference(args):
    """The reference's own CPU path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import cpu_ref

    K, W = args.steps, args.warmup
    res = cpu_ref.time_decode_steps(steps=K, warmup=min(max(W, 1), 2), budget_s=90.0)
    out = {"impl": "reference", "metric": "decode_attn_hbm_gbps", "value": res["value"], "unit": "GB/s",
           "n_gpus": args.gpus, "steps": res["steps"], "warmup": res["warmup"], "ms_per_step": res["ms_per_step"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": res["dtype"],
           "data": "synthetic",
           "config": {"workload": "C2 Llama-3-8B decode: batch 64/GPU x 4096 ctx, 32q/8kv heads, D128, page16; "
                                  "step = split_rotary+append+decode of one layer", "sample": res["sample"]},
           "cpu_baseline": {"value": res["value"], "unit": "GB/s", "cores": res["cores"], "kind": res["kind"],
                            "sample": res["sample"]},
           "e2e": {"value": res["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--windows", type=int, default=11, help="timed windows of --steps steps each; the median is reported")
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="decode", choices=["decode", "prefill", "append", "c5", "c5decode", "c4"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f16"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and ref_gpu legs")
    ap.add_argument("--no-sub", action="store_true", help="headline only (profiler runs)")
    ap.add_argument("--no-cupti", action="store_true", help="skip the CUPTI kernel list (profiler runs: ncu owns CUPTI)")
    ap.add_argument("--gather", choices=["p2p", "p2p-unfused", "nccl"], default="p2p",
                    help="c4: how the per-head outputs are re-assembled across ranks")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs: launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
