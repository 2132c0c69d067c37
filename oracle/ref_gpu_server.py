"""TEST / BENCH INFRASTRUCTURE ONLY -- runs INSIDE THE REFERENCE'S OWN RUNTIME on the GPU box, as a separate process.

The process loads the unmodified reference's runtime libraries (oracle/_ref/tvm_cuda/lib: libtvm_ffi 0.1.14, libtvm_runtime,
libtvm_runtime_cuda, libtvm_runtime_extra -- built from /root/reference by oracle/ref_harness/build_tvm_cuda.sh) and its
vendored tvm-ffi Python package; the product's own process uses pip tvm-ffi 0.1.9, and two libtvm_ffi versions must not
share a process.  Never imported by tvm_b200/; started by oracle/ref_gpu.py from tests/ and bench.py.  Two jobs:

  route_a   INTEGRATION.md route A, on a GPU: the reference's C++ PagedAttentionKVCacheObj
            (`vm.builtin.paged_attention_kv_cache_create`, src/runtime/vm/paged_kv_cache.cc:2535-2639) constructed with
            tvm_b200's 13 packed callbacks (tvm_ffi.load_module(libtvm_b200.so)) in place of its TIR kernels, driven
            through the scenario programs of tests/golden/kvcache_*.npz -- the op lists of the reference's own scenario
            tests (test_runtime_builtin_paged_attention_kv_cache_{cpu,tir}.py) plus the randomised ones.  Every attention
            output and debug_get_kv dump is compared with what the reference produced with its own kernels (recorded in
            the fixture).  The cache object, its aux-array manager (byte_offset views of one merged buffer), its stream
            switching and its callback adapters (attn_backend.h) are all the reference's code.
  kernels   the reference's own GPU TIR kernels (oracle/_ref/ref_gpu_kernels_<dtype>_*.so, emitted by
            oracle/ref_harness/emit_ref_gpu_kernels.py) on tensors given in an .npz: outputs back in an .npz, optional
            CUDA-event timing.  This is what the 2e-3 / 1e-2 tolerance is defined against, fp16 and bf16.

Protocol: one JSON document on stdout, last line.  Exit code 0 also when a comparison fails (the JSON says so)."""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys
import traceback
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
REF = HERE / "_ref" / "tvm_cuda"


def _boot():
    """Import the reference's tvm-ffi (not pip's) and load its runtime libraries."""
    sys.path.insert(0, str(REF / "py"))
    import tvm_ffi

    assert tvm_ffi.__version__.startswith("0.1.14"), f"wrong tvm_ffi {tvm_ffi.__version__} from {tvm_ffi.__file__}"
    for lib in ("libtvm_runtime.so", "libtvm_runtime_cuda.so", "libtvm_runtime_extra.so"):
        ctypes.CDLL(str(REF / "lib" / lib), mode=ctypes.RTLD_GLOBAL)
    return tvm_ffi


def _torch_dtype(torch, name):
    return {"float16": torch.float16, "bfloat16": torch.bfloat16, "float32": torch.float32, "int32": torch.int32}[name]


def _to_dev(torch, arr, dtype):
    """numpy -> cuda tensor; 16-bit floats travel as uint16 bit patterns."""
    import numpy as np

    if dtype in ("float16", "bfloat16") and arr.dtype == np.uint16:
        return torch.from_numpy(arr.view(np.int16).copy()).view(_torch_dtype(torch, dtype)).cuda()
    return torch.from_numpy(np.ascontiguousarray(arr)).to(_torch_dtype(torch, dtype)).cuda()


def _bits(torch, t):
    import numpy as np

    if t.dtype in (torch.float16, torch.bfloat16):
        return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


# ---------------------------------------------------------------------------------------------------------------------
def run_kernels(spec):
    import numpy as np
    import torch

    tvm_ffi = _boot()
    mod = tvm_ffi.load_module(str(HERE / "_ref" / spec["module"]))
    z = np.load(spec["inputs"]) if spec.get("inputs") else {}
    T = {}
    for name, d in spec["tensors"].items():
        if d.get("init", "npz") == "npz":
            T[name] = _to_dev(torch, z[name], d["dtype"])
        elif d["init"] == "randn":  # bench-sized inputs are generated on the device (a C2 cache is 1 GiB)
            g = torch.Generator(device="cuda")
            g.manual_seed(int(d.get("seed", 0)))
            T[name] = torch.randn(tuple(d["shape"]), generator=g, device="cuda", dtype=_torch_dtype(torch, d["dtype"]))
        else:
            T[name] = torch.zeros(tuple(d["shape"]), dtype=_torch_dtype(torch, d["dtype"]), device="cuda")
    F = {name: tvm_ffi.from_dlpack(t) for name, t in T.items()}

    def call(c):
        mod[c["fn"]](*[F[a] if isinstance(a, str) else a for a in c["args"]])

    for c in spec["calls"]:
        call(c)
    torch.cuda.synchronize()
    out = {"ok": True, "timings_ms": {}}
    np.savez(spec["outputs"], **{n: _bits(torch, T[n]) for n in spec["fetch"]})
    for tm in spec.get("time", []):
        c = spec["calls"][tm["call"]]
        for _ in range(tm.get("warmup", 3)):
            call(c)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if tm.get("flush_l2") else None
        times = []
        for _ in range(tm.get("iters", 10)):
            if flush is not None:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call(c)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        out["timings_ms"][c["fn"]] = {"median": times[len(times) // 2], "min": times[0], "iters": len(times)}
    return out


# ---------------------------------------------------------------------------------------------------------------------
CB_ORDER = ["transpose_append", "prefill_ragged", "prefill", "decode", "prefill_sliding_window", "decode_sliding_window",
            "tree_paged", "tree_ragged", "merge", "split_rotary", "copy_single_page", "debug_get_kv", "compact_copy"]
CB_NAMES = {"transpose_append": "f_transpose_append", "prefill_ragged": "f_attention_prefill_ragged",
            "prefill": "f_attention_prefill", "decode": "f_attention_decode",
            "prefill_sliding_window": "f_attention_prefill_sliding_window",
            "decode_sliding_window": "f_attention_decode_sliding_window",
            "tree_paged": "f_attention_prefill_with_tree_mask_paged_kv", "tree_ragged": "f_attention_prefill_with_tree_mask",
            "merge": "f_merge_inplace", "split_rotary": "f_split_rotary", "copy_single_page": "f_copy_single_page",
            "debug_get_kv": "f_debug_get_kv", "compact_copy": "f_compact_copy"}


class RouteA:
    """The reference's cache object around tvm_b200's callbacks (one kernel-set context per cache: the fixture's
    rope_theta / rope_scale / layer window are what the reference would have compiled into its PrimFuncs)."""

    def __init__(self, tvm_ffi, torch, b200, cfg):
        self.tvm_ffi, self.torch, self.cfg = tvm_ffi, torch, cfg
        g = tvm_ffi.get_global_func
        self.ctx = b200["context_create"]()
        self._release = b200["context_release"]
        bind = b200["bind_context"]
        bind(self.ctx, "set_rope_params")(float(cfg["rope_theta"]), float(cfg["rope_scale"]), 0)
        bind(self.ctx, "set_layer_sliding_window_size")(int(cfg.get("layer_sliding_window_size") or 1024))
        bind(self.ctx, "set_rope_scaling")(0, 1.0, 0.0, 0.0, 0.0)
        self.calls = {n: 0 for n in CB_ORDER}
        fns = {}
        for n in CB_ORDER:
            fns[n] = self._counted(n, bind(self.ctx, CB_NAMES[n]))
        self.f = {n: g("vm.builtin." + n) for n in [
            "kv_state_clear", "kv_state_add_sequence", "kv_state_remove_sequence", "kv_state_fork_sequence",
            "kv_state_popn", "kv_state_begin_forward", "kv_state_end_forward",
            "attention_kv_cache_enable_sliding_window_for_seq", "attention_kv_cache_commit_accepted_token_tree_nodes",
            "attention_kv_cache_attention_with_fused_qkv", "attention_kv_cache_empty",
            "attention_kv_cache_get_num_available_pages", "attention_kv_cache_get_total_sequence_length",
            "attention_kv_cache_debug_get_kv", "attention_kv_cache_self_attention",
            "attention_kv_cache_cross_attention", "attention_kv_cache_attention_with_shared_kv",
            "attention_kv_cache_merge_attn_output_inplace"]}
        S = tvm_ffi.Shape
        cache_config = [cfg["reserved_nseq"], cfg["max_total_seq"], cfg["prefill_chunk"], cfg["page_size"],
                        int(cfg.get("support_sliding_window", 0))]
        if cfg.get("layer_sliding_window_size") is not None:
            cache_config.append(cfg["layer_sliding_window_size"])
        lb = cfg.get("layer_begin", 0)
        kinds = cfg.get("attn_kinds") or [0] * (lb + cfg["num_layers"])
        self.tdt = _torch_dtype(torch, cfg["dtype"])
        init = tvm_ffi.from_dlpack(torch.empty((1,), dtype=self.tdt, device="cuda"))
        self.cache = g("vm.builtin.paged_attention_kv_cache_create")(
            S(cache_config), S([lb, lb + cfg["num_layers"]]), cfg["num_qo_heads"], cfg["num_kv_heads"], cfg["head_dim"],
            cfg["head_dim"], S(kinds), False, int(cfg["rope_mode"]), float(cfg["rope_scale"]), float(cfg["rope_theta"]),
            None, init, fns["transpose_append"], None, ["tirx", fns["prefill_ragged"]], ["tirx", fns["prefill"]],
            ["tirx", fns["decode"]], ["tirx", fns["prefill_sliding_window"]], ["tirx", fns["decode_sliding_window"]],
            ["tirx", fns["tree_paged"]], ["tirx", fns["tree_ragged"]], [], [fns["merge"], fns["merge"]],
            fns["split_rotary"], fns["copy_single_page"], fns["debug_get_kv"], fns["compact_copy"])

    def _counted(self, name, fn):
        if os.environ.get("ROUTE_A_COUNT", "1") == "0":
            return fn
        calls = self.calls

        def counted(*args):
            calls[name] += 1
            return fn(*args)

        return self.tvm_ffi.convert(counted)

    def call(self, name, *args):
        return self.f[name](self.cache, *args)

    def close(self):
        self.cache = None
        self._release(self.ctx)


def _close(np, name, got, want, atol=2e-3, rtol=1e-2):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want)
    bad = err > atol + rtol * np.abs(want)
    if bad.any():
        i = np.unravel_index(err.argmax(), err.shape)
        raise AssertionError(f"{name}: {int(bad.sum())}/{bad.size} out of tolerance, max abs err {err.max():.3e} at {i} "
                             f"(got {got[i]}, want {want[i]})")
    return float(err.max()) if err.size else 0.0


def replay_fixture(tvm_ffi, torch, b200, name):
    import numpy as np

    z = np.load(ROOT / "tests" / "golden" / f"kvcache_{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    cfg = meta["config"]
    S = tvm_ffi.Shape
    ra = RouteA(tvm_ffi, torch, b200, cfg)
    L, hq, hkv, d = cfg["num_layers"], cfg["num_qo_heads"], cfg["num_kv_heads"], cfg["head_dim"]
    lb = cfg.get("layer_begin", 0)
    sm = d ** -0.5
    dev = lambda a: tvm_ffi.from_dlpack(a)  # noqa: E731
    stats = {"name": name, "forwards": 0, "dumps": 0, "max_err_o": 0.0, "max_err_k": 0.0}

    def gen(seed, shape):
        rng = np.random.default_rng(seed)
        return rng.random(shape, dtype=np.float32).astype(cfg["dtype"])

    for idx, (op, res) in enumerate(zip(meta["ops"], meta["results"])):
        k = op["op"]
        where = f"{name} op {idx} {k}"
        if k == "clear":
            ra.call("kv_state_clear")
        elif k == "add":
            ra.call("kv_state_add_sequence", op["seq"])
        elif k == "remove":
            ra.call("kv_state_remove_sequence", op["seq"])
        elif k == "fork":
            ra.call("kv_state_fork_sequence", op["parent"], op["child"], op["pos"])
        elif k == "popn":
            ra.call("kv_state_popn", op["seq"], op["n"])
        elif k == "enable_sw":
            ra.call("attention_kv_cache_enable_sliding_window_for_seq", op["seq"], op["window"], op["sink"])
        elif k == "commit":
            ra.call("attention_kv_cache_commit_accepted_token_tree_nodes", S(op["seq_ids"]), S(op["leaves"]))
        elif k == "query":
            assert bool(ra.call("attention_kv_cache_empty")) == res["empty"], where
            assert int(ra.call("attention_kv_cache_get_num_available_pages")) == res["num_available_pages"], where
            assert int(ra.call("attention_kv_cache_get_total_sequence_length")) == res["total_sequence_length"], where
        elif k == "debug_get_kv":
            n = op["end"] - op["start"]
            kk = torch.zeros((L, n, hkv, d), dtype=ra.tdt, device="cuda")
            vv = torch.zeros_like(kk)
            ra.call("attention_kv_cache_debug_get_kv", op["seq"], op["start"], op["end"], dev(kk), dev(vv))
            torch.cuda.synchronize()
            # V is a pure copy of the inputs: bit-exact; K went through RoPE (fp32 trig, cast to dtype)
            assert np.array_equal(vv.float().cpu().numpy(), np.asarray(z[f"v_{idx}"], np.float32)), f"{where}: V differs"
            stats["max_err_k"] = max(stats["max_err_k"], _close(np, f"{where} K", kk.float().cpu().numpy(), z[f"k_{idx}"],
                                                                atol=4e-3))
            stats["dumps"] += 1
        elif k == "debug_get_kv_rejected":
            try:
                kk = torch.zeros((L, 1, hkv, d), dtype=ra.tdt, device="cuda")
                ra.call("attention_kv_cache_debug_get_kv", op["seq"], 0, 1, dev(kk), dev(kk))
                raise AssertionError(f"{where}: the reference did not refuse")
            except Exception as e:  # noqa: BLE001
                assert "Only MHA" in str(e), f"{where}: {e}"
        elif k == "forward":
            tree = S(op["tree"]) if op["tree"] is not None else None
            if tree is None:
                ra.call("kv_state_begin_forward", S(op["seq_ids"]), S(op["lens"]))
            else:
                ra.call("kv_state_begin_forward", S(op["seq_ids"]), S(op["lens"]), tree)
            n = sum(op["lens"])
            qkv = gen(op["seed"], (L, n, hq + 2 * hkv, d))
            shared = bool(op.get("shared"))
            q2 = gen(op["seed"] + 500000, (L, n, hq, d)) if shared else None
            outs, souts = [], []
            for layer in range(L):
                tq = torch.from_numpy(qkv[layer]).cuda()
                o = torch.full((n, hq, d), float("nan"), dtype=tq.dtype, device="cuda")
                ra.call("attention_kv_cache_attention_with_fused_qkv", lb + layer, sm, dev(tq), dev(o))
                outs.append(o)
                if shared:
                    o2 = torch.full((n, hq, d), float("nan"), dtype=tq.dtype, device="cuda")
                    ra.call("attention_kv_cache_attention_with_shared_kv", lb + layer, sm, dev(torch.from_numpy(q2[layer]).cuda()),
                            dev(tq[:, hq:hq + hkv].contiguous()), dev(tq[:, hq + hkv:].contiguous()), dev(o2))
                    souts.append(o2)
            ra.call("kv_state_end_forward")
            torch.cuda.synchronize()
            got = np.stack([o.float().cpu().numpy() for o in outs])
            stats["max_err_o"] = max(stats["max_err_o"], _close(np, f"{where} O", got, z[f"o_{idx}"]))
            if shared:
                gs = np.stack([o.float().cpu().numpy() for o in souts])
                stats["max_err_o"] = max(stats["max_err_o"], _close(np, f"{where} O(shared kv)", gs, z[f"os_{idx}"]))
            assert int(ra.call("attention_kv_cache_get_num_available_pages")) == res["num_available_pages"], where
            stats["forwards"] += 1
        elif k == "forward_split":
            ra.call("kv_state_begin_forward", S(op["seq_ids"]), S(op["lens"]))
            n = sum(op["lens"])
            qkv = gen(op["seed"], (L, n, hq + 2 * hkv, d))
            for layer in range(L):
                tq = torch.from_numpy(qkv[layer]).cuda()
                q, kk, vv = tq[:, :hq].contiguous(), tq[:, hq:hq + hkv].contiguous(), tq[:, hq + hkv:].contiguous()
                o_self = torch.zeros((n, hq, d), dtype=tq.dtype, device="cuda")
                lse_self = torch.full((n, hq), -5e4, dtype=torch.float32, device="cuda")
                o_cross, lse_cross = torch.zeros_like(o_self), torch.full_like(lse_self, -5e4)
                ra.call("attention_kv_cache_self_attention", lb + layer, sm, dev(q), dev(kk), dev(vv), dev(o_self), dev(lse_self))
                ra.call("attention_kv_cache_cross_attention", lb + layer, sm, dev(q), dev(o_cross), dev(lse_cross))
                ra.call("attention_kv_cache_merge_attn_output_inplace", dev(o_self), dev(lse_self), dev(o_cross), dev(lse_cross))
                torch.cuda.synchronize()
                stats["max_err_o"] = max(stats["max_err_o"],
                                         _close(np, f"{where} O", o_self.float().cpu().numpy(), z[f"o_{idx}"][layer]),
                                         _close(np, f"{where} LSE", lse_self.cpu().numpy(), z[f"lse_{idx}"][layer]))
            ra.call("kv_state_end_forward")
            stats["forwards"] += 1
        else:
            raise ValueError(k)
    stats["callbacks"] = dict(ra.calls)
    ra.close()
    return stats


def run_route_a(names):
    import torch

    tvm_ffi = _boot()
    b200 = tvm_ffi.load_module(str(ROOT / "tvm_b200" / "lib" / "libtvm_b200.so"))
    launches0 = int(b200["launch_count"]())
    results, ok = [], True
    for name in names:
        try:
            results.append(dict(replay_fixture(tvm_ffi, torch, b200, name), ok=True))
        except Exception as e:  # noqa: BLE001
            ok = False
            results.append({"name": name, "ok": False, "error": f"{type(e).__name__}: {e}",
                            "trace": traceback.format_exc().splitlines()[-6:]})
    return {"ok": ok, "tvm_ffi": tvm_ffi.__version__, "device": torch.cuda.get_device_name(0),
            "kernel_launches": int(b200["launch_count"]()) - launches0, "fixtures": results}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("job", choices=["route_a", "kernels"])
    ap.add_argument("--fixtures", default="")
    ap.add_argument("--spec", default="")
    a = ap.parse_args()
    try:
        if a.job == "route_a":
            out = run_route_a([n for n in a.fixtures.split(",") if n])
        else:
            out = run_kernels(json.loads(Path(a.spec).read_text()))
    except Exception as e:  # noqa: BLE001
        out = {"ok": False, "error": f"{type(e).__name__}: {e}", "trace": traceback.format_exc().splitlines()[-12:]}
    print(json.dumps(out))
    sys.stdout.flush()
    os._exit(0)  # skip interpreter teardown: two runtimes' static destructors race with torch's CUDA context


if __name__ == "__main__":
    main()
