#!/usr/bin/env python
"""bench.py -- PagedKVCache attention hot path on B200 (contract: see DESIGN.md "Measurement").

A step = one layer-call of the decode hot path over one batch, exactly the callback sequence the
reference cache issues per layer for a plain decode step (paged_kv_cache.cc:1355-1401):
    f_split_rotary -> f_transpose_append -> f_attention_decode
on BASELINE.json configs[1] (C2: Llama-3-8B shape, batch 64 decode at 4K context, bf16 paged KV).
With --gpus N the sequence batch is split across ranks (64 sequences per rank, global batch 64*N, weak
scaling) and the per-rank outputs are re-assembled with an NCCL all-gather inside the timed region.

    metric  decode_attn_hbm_gbps = algorithmic bytes of the step (BASELINE.md section 3 formulas) / time
    value   device-resident inputs (CUDA events, max over ranks)
    e2e     the same through the tvm-ffi packed functions with HOST inputs: pinned qkv + merged aux
            arrays copied host->device and O copied device->host every step, inside the timed region
    roofline  the decode kernel alone against the measured HBM copy peak (MEASURED_PEAKS.json)
    cpu_baseline  the reference's CPU path (oracle/_ref if built, else the NumPy oracle port) on a slice

`--impl reference` times the CPU path only (rank 0) and prints the same JSON line shape.
`--workload prefill` reports C3 (ragged causal prefill 16x2048) TFLOP/s instead.
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
# from the committed `ncu --set full` captures under profiles/ (a profiler number: reported next to the algorithmic
# bytes, never used for timing)
NCU_DRAM_BYTES = {"decode_c2": 1074640000 + 6002176, "prefill_c3": 402785280 + 220196864}


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
# workload C2: decode
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
        e = 2
        return (self.B * self.L * self.Hkv * self.D * 2 * e + 2 * self.B * self.Hq * self.D * e + 4 * self.B * self.Hq
                + 4 * (self.nnz + 4 * self.B + 1))

    def append_bytes(self):
        return self.B * self.Hkv * self.D * 2 * 2 * 2 + 4 * self.B

    def rotary_bytes(self):
        return 2 * self.B * (self.Hq + 2 * self.Hkv) * self.D * 2 + 4 * self.B

    def step_bytes(self):
        return self.decode_bytes() + self.append_bytes() + self.rotary_bytes()

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
    """C3: 16 sequences x 2048 new tokens, empty cache -> ragged causal prefill (+ rotary + append)."""

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
        return 4 * self.D * self.Hq * self.nseq * (self.L * (self.L + 1) // 2)

    def run(self, capi):
        capi.attention_prefill_ragged(self.q, self.indptr, self.k, self.v, self.indptr, self.qpos, self.kofs, self.o,
                                      self.lse, 1, 0, 1.0, 5e5, self.sm_scale)


# ---------------------------------------------------------------------------------------------------
# clocks sampling (NVML) during the timed region
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index=0, period_s=0.002):
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
# CPU baseline (the reference's CPU path): oracle/_ref if present, else the NumPy oracle port
# ---------------------------------------------------------------------------------------------------
def cpu_decode_baseline(L=4096, Hq=32, Hkv=8, D=128, B=1, repeats=50):
    """Times the CPU decode kernel on a bounded slice (B sequences of the C2 shape, fp16 like the
    reference's CPU tests).  Returns dict(value GB/s, unit, cores, kind, sample)."""
    from oracle import cpu_ref

    return cpu_ref.time_decode(B=B, L=L, Hq=Hq, Hkv=Hkv, D=D, repeats=repeats)


# ---------------------------------------------------------------------------------------------------
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


def run_own(args):
    import torch

    from tvm_b200 import capi

    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    capi.lib()
    peaks, peak_src = measured_peaks()
    if args.workload == "prefill":
        return run_prefill(args, capi, rank, world, dev, peaks, peak_src)
    if args.workload == "c4":
        return run_c4(args, capi, rank, world, dev, peaks, peak_src)
    w = DecodeWorkload(seed=rank, device=dev)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: F811

    # Batch split (weak scaling): every rank owns 64 whole sequences -- page table, KV pages, queries and outputs -- so
    # the path has NO exchange step (a sequence's attention output feeds that rank's own next layer); NCCL is only
    # used for the barrier / max-over-ranks timing.  The head-sharded mode (--workload c4) is the one with a real
    # exchange (per-head outputs are re-assembled) and keeps its all-gather inside the timed step.
    def step():
        w.run_step_fused(capi)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    K = args.steps
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = capi.launch_count()
    # pass 1 -- the timed region of `value`: exactly K steps, nothing but the step's own launches on the stream (an
    # event recorded between the kernels would break their programmatic dependent launch and cost ~10 us per step)
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        e_beg.record()
        for _ in range(K):
            step()
        e_end.record()
        torch.cuda.synchronize()
    launches = capi.launch_count() - n0
    if world > 1:
        dist.barrier()
    total_ms = e_beg.elapsed_time(e_end)
    # pass 2 -- roofline of the dominant kernel (the fused decode launch + its split-KV merge): the same K steps with
    # CUDA events around every launch pair
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K)]
    for i in range(K):
        ev[2 * i].record()
        w.run_step_fused(capi)
        ev[2 * i + 1].record()
    torch.cuda.synchronize()
    decode_ms = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(K)) / K
    if world > 1:
        t = torch.tensor([total_ms, decode_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, decode_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / K
    value = world * w.step_bytes() / (ms_per_step * 1e-3) / 1e9
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))
    dec_gbs = w.step_bytes() / (decode_ms * 1e-3) / 1e9
    out = {
        "metric": "decode_attn_hbm_gbps", "value": round(value, 1), "unit": "GB/s", "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 5), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "C2 Llama-3-8B decode: batch 64/GPU x 4096 ctx, 32q/8kv heads, D128, page16, bf16 "
                               "paged KV; step = split_rotary+append+decode of one layer (one fused launch + the split-KV merge)",
                   "global_batch": w.B * world, "seq_len": w.L, "parallelism": f"batch-split x{world} (no data-path collective)",
                   "l2": "KV working set 1 GiB/GPU > 126 MB L2 (no flush needed)"},
        "tok_s_layer": round(world * w.B / (ms_per_step * 1e-3), 1),
        "roofline": {"bound": "hbm", "kernel": "decode_kernel<FUSED qkv>(+decode_merge_kernel)", "achieved": round(dec_gbs, 1),
                     "peak": hbm_peak, "unit": "GB/s", "frac": round(dec_gbs / hbm_peak, 4), "traffic": NCU_DRAM_BYTES["decode_c2"],
                     "traffic_source": "profiles/r1_decode_v3_ncu.md: dram read+write of one launch, ncu --set full",
                     "peak_source": f"of {peak_src}", "algorithmic_bytes": w.step_bytes(),
                     "kernel_ms": round(decode_ms, 5), "frac_of_spec_8000": round(dec_gbs / 8000.0, 4)},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    try:
        e2e = {"value": None, "skipped": "--no-e2e"} if args.no_e2e else run_e2e(args, w, world, dist)
    except Exception as e:  # pragma: no cover
        e2e = {"value": None, "error": repr(e)[:300]}
    if rank == 0:
        out["e2e"] = e2e
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
    K = max(10, args.steps // 3)
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


def run_prefill(args, capi, rank, world, dev, peaks, peak_src):
    import torch

    w = PrefillWorkload(seed=rank, device=dev, dtype=args.dtype)
    for _ in range(max(args.warmup, 3)):
        w.run(capi)
    torch.cuda.synchronize()
    K = max(1, min(args.steps, 50))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = capi.launch_count()
    with ClockSampler(dev.index or 0) as clk:
        e0.record()
        for _ in range(K):
            w.run(capi)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    tf = w.flops() / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"]))
    out = {"metric": "prefill_tflops", "value": round(tf * world, 2), "unit": "TFLOP/s", "n_gpus": world, "steps": K,
           "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
           "config": {"workload": "C3 ragged causal prefill 16x2048, 32q/8kv heads, D128, bf16", "l2": "q/k/v/o "
                      "0.67 GB > L2"},
           "roofline": {"bound": "tensor", "achieved": round(tf, 2), "peak": peak, "unit": "TFLOP/s",
                        "frac": round(tf / peak, 4), "traffic": NCU_DRAM_BYTES["prefill_c3"],
                        "traffic_source": "profiles/r1_prefill_tc05_v6_ncu.md: dram read+write of one launch (bytes)",
                        "peak_source": f"of {peak_src} (burst; sustained " + str(peaks.get("bf16_tflops_sustained")) + ")",
                        "frac_of_sustained": (round(tf / float(peaks["bf16_tflops_sustained"]), 4)
                                              if peaks.get("bf16_tflops_sustained") else None),
                        "algorithmic_flops": w.flops()},
           "gpu_launches": int(capi.launch_count() - n0), "clocks": clk.summary()}
    if rank == 0:
        print(json.dumps(out), flush=True)


def run_c4(args, capi, rank, world, dev, peaks, peak_src):
    """C4: Llama-3-70B GQA decode (64 q / 8 kv heads), batch 256 at 8K context, KV-head groups sharded across the ranks
    (strong scaling: total work fixed), per-head outputs re-assembled with one NCCL all-gather per step."""
    import torch

    from tvm_b200 import sharding

    Hq, Hkv, B, L = 64, 8, 256, 8192
    q0, q1, k0, k1 = sharding.head_shard(Hq, Hkv, world, rank)
    w = DecodeWorkload(B=B, L=L, Hq=q1 - q0, Hkv=k1 - k0, seed=0, device=dev)  # same page table on every rank
    gather, gather_kind = None, "none"
    if world > 1:
        import torch.distributed as dist

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

    def step():
        if gather is not None and gather_kind == "p2p-unfused":  # rotary+append launch, then decode + head gather
            w.run_rotary_append(capi)
            return w.run_decode_gather(capi, gather)
        if gather is not None:  # rotary + append + decode + head gather: one fused launch, the merge, the flag wait
            return gather.decode_fused_qkv(capi, w.qkv, w.q_rope_position, w.append_position, w.pages, w.page_indptr,
                                           w.page_values, w.length_info, w.k_rope_pos_offset, w.o, w.lse, 1, w.rope_scale,
                                           w.rope_theta, w.sm_scale)
        w.run_step_fused(capi)
        if world > 1:
            return sharding.all_gather_heads(w.o)
        return w.o

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    K = args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = capi.launch_count()
    with ClockSampler(dev.index or 0) as clk:
        e0.record()
        for _ in range(K):
            step()
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
    total_bytes = world * w.step_bytes()
    gbs = total_bytes / (ms * 1e-3) / 1e9
    hbm_peak = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))
    out = {"metric": "decode_attn_hbm_gbps", "value": round(gbs, 1), "unit": "GB/s", "n_gpus": world, "steps": K,
           "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 5), "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
           "config": {"workload": "C4 Llama-3-70B GQA decode: batch 256 x 8192 ctx, 64q/8kv heads sharded by KV-head group, "
                                  "D128, page16, bf16; step = split_rotary+append+decode+re-assembly of the per-head O on every rank",
                      "head_gather": {"p2p": "in-kernel NVLink peer stores + flags, fused qkv step (tvmb200_attention_decode_fused_qkv_gather)",
                                      "p2p-unfused": "in-kernel NVLink peer stores + flags (tvmb200_attention_decode_gather)",
                                      "nccl": "ncclAllGather behind the kernel", "none": "single GPU"}[gather_kind],
                      "global_batch": B, "seq_len": L, "parallelism": f"tp{world} (KV-head groups)",
                      "l2": "KV working set >= 1 GiB/GPU > 126 MB L2"},
           "tok_s_layer": round(B / (ms * 1e-3), 1),
           "roofline": {"bound": "hbm", "achieved": round(gbs / world, 1), "peak": hbm_peak, "unit": "GB/s",
                        "frac": round(gbs / world / hbm_peak, 4), "traffic": None, "peak_source": f"of {peak_src}",
                        "note": "whole step per GPU incl. all-gather"},
           "gpu_launches": int(capi.launch_count() - n0), "clocks": clk.summary()}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="decode", choices=["decode", "prefill", "c4"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f16"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", choices=["p2p", "p2p-unfused", "nccl"], default="p2p",
                    help="c4 workload: how the per-head outputs are re-assembled across ranks")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs: launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
