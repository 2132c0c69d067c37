"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference:
its C++ PagedAttentionKVCacheObj (src/runtime/vm/paged_kv_cache.cc) over its own CPU TIR kernels.

    bash oracle/ref_harness/build_tvm.sh /tmp/tvm_ref          # once, ~25 min
    source /tmp/tvm_ref/env.sh && python oracle/ref_harness/gen_golden.py

Each fixture holds a scenario "program" (the op list of one of the reference's own scenario tests,
tests/python/relax/test_runtime_builtin_paged_attention_kv_cache_cpu.py:706-1104), and for every op what the
reference did: the exact callback sequence with every int32 aux array and scalar it passed (bit-exact targets for
our host cache), the attention outputs, and debug_get_kv dumps.  Inputs are regenerated from the recorded seeds
(`qkv_for`), so only outputs are stored.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
OUT = os.path.join(HERE, "..", "..", "tests", "golden")


def qkv_for(seed, num_layers, n, hq, hkv, d, dtype="float16"):
    """The step's fused qkv input, [num_layers, n, hq + 2 hkv, d], U[0,1) like the reference tests."""
    rng = np.random.default_rng(seed)
    return rng.random((num_layers, n, hq + 2 * hkv, d), dtype=np.float32).astype(dtype)


# ---- programs (op lists) -------------------------------------------------------------------------------------------
class Program:
    def __init__(self):
        self.ops = []
        self.known = set()
        self.seed = 1000

    def clear(self):
        self.ops.append({"op": "clear"})
        self.known = set()

    def forward_split(self, batch):
        """begin_forward + per layer self_attention / cross_attention / merge_attn_output_inplace on the batch's q, k, v
        (kv_state.cc:84-115), then popn of the reserved lengths: these entries append nothing, so the reservation is
        rolled back and the cache content stays defined.  Lengths >= 2 keep the batch a prefill (append after attention),
        so the cross part only reads tokens that were really appended."""
        assert all(ln >= 2 for _, ln in batch)
        self.seed += 1
        self.ops.append({"op": "forward_split", "seq_ids": [s for s, _ in batch], "lens": [ln for _, ln in batch],
                         "seed": self.seed})
        for s, ln in batch:
            self.ops.append({"op": "popn", "seq": s, "n": ln})

    def forward(self, batch, trees=None, leaves=None, shared=False):
        """batch items: (seq_id, len) or ((seq_id, parent, fork_pos), len) -- the reference's apply_attention.
        shared: every layer's attention_with_fused_qkv is followed by attention_with_shared_kv on the same layer with a
        second query (test_attention_with_shared_kv, ..._cpu.py:654-703)."""
        seq_ids, lens = [], []
        for item, ln in batch:
            if isinstance(item, tuple):
                sid, parent, pos = item
                self.ops.append({"op": "fork", "parent": parent, "child": sid, "pos": pos})
                self.known.add(sid)
            else:
                sid = item
                if sid not in self.known:
                    self.ops.append({"op": "add", "seq": sid})
                    self.known.add(sid)
            seq_ids.append(sid)
            lens.append(ln)
        flat = None
        if trees is not None:
            flat = []
            for t, ln in zip(trees, lens):
                flat += t[-ln:]
        self.seed += 1
        self.ops.append({"op": "forward", "seq_ids": seq_ids, "lens": lens, "tree": flat, "seed": self.seed})
        if shared:
            self.ops[-1]["shared"] = True
        if leaves is not None:
            self.ops.append({"op": "commit", "seq_ids": seq_ids, "leaves": leaves})

    def op(self, **kw):
        self.ops.append(kw)
        if kw["op"] == "add":
            self.known.add(kw["seq"])
        if kw["op"] == "fork":
            self.known.add(kw["child"])
        if kw["op"] == "remove":
            self.known.discard(kw["seq"])

    def dump_all(self, lengths):
        for sid, ln in lengths.items():
            if ln > 0:
                self.ops.append({"op": "debug_get_kv", "seq": sid, "start": 0, "end": ln})


def prog_prefill_and_decode():
    p = Program()
    seqs = [[(0, 6)], [(1, 8)], [(2, 11)], [(3, 16)], [(4, 19), (5, 20)], [(6, 21), (7, 24)],
            [(2, 5), (4, 7), (8, 24)], [(6, 13)], [(8, 19)], [(0, 1)], [(1, 3), (3, 8), (5, 12), (7, 11)]]
    seqs += [[(i, 1) for i in range(9)]] * 2 + [[(0, 1), (2, 1), (4, 1), (6, 1), (8, 1)], [(4, 1), (5, 1), (6, 1), (7, 1), (8, 1)]]
    ln = {}
    for b in seqs:
        p.forward(list(b))
        for s, n in b:
            ln[s] = ln.get(s, 0) + n
    p.dump_all(ln)
    return p


def prog_remove_and_popn():
    p = Program()
    p.forward([(0, 35), (1, 88), (2, 17), (3, 4)])
    p.forward([((4, 3, -1), 35)])
    ln = {0: 35, 1: 88, 2: 17, 3: 4, 4: 39}
    for sid, n in [(0, 17), (1, 57), (2, 16), (3, 0), (4, 37)]:
        p.op(op="popn", seq=sid, n=n)
        ln[sid] -= n
    p.dump_all(ln)
    p.forward([(0, 3), (1, 1), (2, 20), (4, 1)])
    p.op(op="remove", seq=1)
    p.forward([(0, 1), (2, 1), (3, 1), (4, 1)])
    for s in (0, 2, 3, 4):
        p.op(op="remove", seq=s)
    p.op(op="query")
    return p


def prog_fork():
    p = Program()
    p.forward([(0, 60), (1, 88), (2, 17), (3, 4)])
    p.forward([((4, 3, -1), 35)])
    p.forward([((5, 0, -1), 20)])
    p.forward([((6, 5, -1), 102)])
    p.forward([((7, 0, -1), 3)])
    p.forward([((8, 5, -1), 71), ((9, 5, -1), 20)])
    for b in [[(2, 1), (4, 1), (7, 1), (6, 1), (8, 1), (9, 1)], [(7, 1), (6, 1), (8, 1), (9, 1)],
              [(7, 1), (1, 1), (6, 1), (2, 1), (8, 1), (4, 1), (9, 1)], [(7, 10), (6, 2), (8, 3), (9, 4)]]:
        p.forward(b)
    p.forward([((10, 1, 33), 11)])
    p.forward([((11, 0, 60), 45), ((12, 0, 15), 14)])
    p.forward([((13, 0, 16), 19), ((14, 0, 17), 19)])
    p.forward([((15, 5, 60), 8), ((16, 5, 80), 10)])
    p.forward([((17, 5, 75), 11), ((18, 5, 76), 45), ((19, 5, 77), 14)])
    for b in [[(6, 1), (11, 1), (13, 1), (9, 1)], [(10, 1), (16, 1), (18, 1), (19, 1)],
              [(8, 1), (15, 1), (17, 1), (12, 1), (14, 1)], [(10, 10), (6, 2), (8, 3), (19, 4)]]:
        p.forward(b)
    p.op(op="debug_get_kv", seq=19, start=0, end=77 + 14 + 1 + 4)
    p.op(op="debug_get_kv", seq=13, start=0, end=16 + 19 + 1)
    for i in range(20):
        p.op(op="remove", seq=i)
    p.op(op="query")
    # fork after page recycle
    p.forward([(0, 7), (1, 24)])
    p.forward([((2, 1, -1), 10)])
    p.forward([((3, 0, -1), 20)])
    p.forward([(2, 1), (3, 1)])
    p.forward([(10, 7), (11, 24)])
    p.forward([((12, 11, -1), 200)])
    p.forward([(10, 1), (12, 1)])
    return p


def prog_unlimited_depth():
    p = Program()
    p.forward([(0, 30)])
    for child, parent, n in [(1, 0, 15), (2, 1, 5), (3, 2, 20), (4, 3, 26), (5, 3, 18), (6, 5, 22), (7, 5, 12),
                             (8, 7, 29), (9, 7, 9), (10, 9, 31), (11, 9, 4)]:
        p.forward([((child, parent, -1), n)])
    for b in [[(3, 1), (6, 1), (9, 1)], [(4, 1), (8, 1), (10, 1)], [(5, 1), (7, 1), (11, 1)]]:
        p.forward(b)
    p.op(op="debug_get_kv", seq=11, start=0, end=30 + 15 + 5 + 20 + 18 + 12 + 9 + 4 + 1)
    for i in range(12):
        p.op(op="remove", seq=i)
    p.op(op="query")
    return p


def prog_sliding_window():
    p = Program()
    sw, sink = [20, 25, 30, 35, 40], [6, 4, 8, 3, 7]
    for sid, (w, s) in enumerate(zip(sw, sink)):
        p.op(op="add", seq=sid)
        p.op(op="enable_sw", seq=sid, window=w, sink=s)
    for b in [[(0, 4)], [(1, 6)], [(2, 6), (3, 7), (4, 7)], [(0, 20), (1, 19), (2, 30), (3, 35), (4, 40)],
              [(0, 6), (1, 5), (2, 4), (3, 3), (4, 2)]]:
        p.forward(b)
    for _ in range(6):
        p.forward([(i, 1) for i in range(5)])
    for sid, w in enumerate(sw):
        p.op(op="debug_get_kv", seq=sid, start=0, end=w)
    return p


def prog_sliding_window_fork():
    p = Program()
    for sid, (w, s) in enumerate(zip([30, 35, 40], [15, 20, 25])):
        p.op(op="add", seq=sid)
        p.op(op="enable_sw", seq=sid, window=w, sink=s)
    p.forward([(0, 12), (1, 18), (2, 28)])
    p.forward([((3, 0, 10), 8), ((4, 1, -1), 20), ((5, 2, 18), 18)])
    p.forward([(0, 9), (1, 15), (2, 4), (3, 10), (4, 3), (5, 7)])
    p.op(op="fork", parent=3, child=6, pos=18)
    p.op(op="enable_sw", seq=6, window=25, sink=24)
    p.forward([(3, 10), (6, 12)])
    return p


def prog_tree_attn():
    p = Program()
    p.forward([(0, 10), (1, 20), (2, 30), (3, 40)])
    trees = [[-1, 0, 0, 1, 1, 2, 2], [-1, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6], [-1, 0, 1, 2, 3, 4, 5, 6, 7, 8],
             [-1, 0, 0, 1, 1, 2, 2, -1, 7, 7, 8, 8, 9, 9]]
    p.forward([(0, 7), (1, 15), (2, 10), (3, 14)], trees=trees, leaves=[6, 11, 6, 13])
    p.op(op="debug_get_kv", seq=1, start=0, end=20 + 4)
    for _ in range(2):
        p.forward([(i, 1) for i in range(4)])
    # all chains
    p.clear()
    p.forward([(0, 10), (1, 20), (2, 30), (3, 40)])
    chains = [list(range(-1, 6)), list(range(-1, 14)), list(range(-1, 9)), list(range(-1, 13))]
    p.forward([(0, 7), (1, 15), (2, 10), (3, 14)], trees=chains, leaves=[2, 6, -1, 4])
    p.forward([(i, 1) for i in range(4)])
    # multi-round trees over cached kv
    p.clear()
    p.forward([(0, 10), (1, 20), (2, 30), (3, 40)])
    for i in range(5):
        nleaf = 2 ** i
        parent = [(k - 1) // 2 for k in range(0, 2 * nleaf - 1)]
        p.forward([(s, nleaf) for s in range(4)], trees=[parent] * 4, leaves=None if i != 4 else [2, 6, -1, 4])
    p.op(op="debug_get_kv", seq=1, start=0, end=20 + 3)
    p.forward([(i, 1) for i in range(4)])
    return p


def prog_random(seed, shared=False, split=False, deep_popn=False):
    """Randomised differential scenario: a few hundred cache operations drawn at random -- prefill chunks, decode
    steps over random subsets, forks at random positions (also of forked sequences), popn (kept inside a sequence's own
    last block, as the reference requires), removals -- so that the host cache's page allocation, block tree, copy-on-
    fork and aux-array construction are compared with the reference far off the hand-written paths.
    deep_popn: popn may cut into blocks shared with other sequences.  shared: every forward also issues
    attention_with_shared_kv per layer; split: self_attention / cross_attention /
    merge_attn_output_inplace steps (rolled back by popn) are mixed in."""
    rng = np.random.default_rng(seed)
    p = Program()
    if shared:
        fwd = p.forward
        p.forward = lambda batch, **kw: fwd(batch, shared=True, **kw)
    length, tail = {}, {}     # seq -> total length, tokens appended since its last fork point / creation
    next_id = 0
    total = 0
    cap = 1800                # max_total_seq is 2048
    for _ in range(60):
        live = sorted(length)
        r = rng.random()
        if (r < 0.25 or not live) and len(live) < 12 and total < cap - 80:
            n = int(rng.integers(1, 70))
            p.forward([(next_id, n)])
            length[next_id], tail[next_id] = n, n
            next_id += 1
            total += n
        elif r < 0.40 and live and len(live) < 12 and total < cap - 80:
            parent = int(rng.choice(live))
            pos = int(rng.integers(1, length[parent] + 1)) if rng.random() < 0.7 else -1
            n = int(rng.integers(1, 40))
            p.forward([((next_id, parent, pos), n)])
            base = length[parent] if pos == -1 else pos
            length[next_id], tail[next_id] = base + n, n
            tail[parent] = min(tail[parent], length[parent] - base) if pos != -1 else 0
            next_id += 1
            total += n
        elif r < 0.80 and live and total < cap - len(live) * 8:
            k = int(rng.integers(1, len(live) + 1))
            sel = [int(x) for x in rng.choice(live, size=k, replace=False)]
            if rng.random() < 0.7:
                batch = [(s_, 1) for s_ in sel]                       # decode
            else:
                batch = [(s_, int(rng.integers(1, 8))) for s_ in sel]  # short multi-token appends
            p.forward(batch)
            for s_, n in batch:
                length[s_] += n
                tail[s_] += n
                total += n
        elif r < 0.90 and live and deep_popn:
            # pop across fork points: the reference re-creates the sequence as a fork of itself at the shorter length
            # (PopN, paged_kv_cache.cc:803-860)
            s_ = int(rng.choice(live))
            if length[s_] > 1:
                n = int(rng.integers(1, length[s_]))
                p.op(op="popn", seq=s_, n=n)
                length[s_] -= n
                tail[s_] = max(tail[s_] - n, 0)
        elif r < 0.90 and live:
            s_ = int(rng.choice(live))
            if tail[s_] > 1:
                n = int(rng.integers(1, tail[s_]))
                p.op(op="popn", seq=s_, n=n)
                length[s_] -= n
                tail[s_] -= n
        elif live:
            leaves = [s_ for s_ in live]
            s_ = int(rng.choice(leaves))
            p.op(op="remove", seq=s_)
            del length[s_], tail[s_]
        if split and length and rng.random() < 0.2 and total < cap - 200:
            live = sorted(length)
            k = int(rng.integers(1, min(len(live), 4) + 1))
            sel = [int(x) for x in rng.choice(live, size=k, replace=False)]
            p.forward_split([(s_, int(rng.integers(2, 30))) for s_ in sel])
        if rng.random() < 0.15:
            p.op(op="query")
    p.dump_all({s_: min(n, 40) for s_, n in length.items()})
    p.op(op="query")
    return p


def prog_random_tree(seed, forks=False):
    """Randomised speculative-decoding program: prefill / decode steps interleaved with token-tree phases -- 1 to 3 rounds
    of random trees over a random subset of the sequences (later rounds hang new nodes under any earlier node), then a
    commit of a random root-to-node path (or of nothing, -1) per sequence."""
    rng = np.random.default_rng(seed)
    p = Program()
    length = {}
    for sid in range(5):
        n = int(rng.integers(3, 50))
        p.forward([(sid, n)])
        length[sid] = n
    next_id = 5
    for _ in range(14):
        live = sorted(length)
        if forks and rng.random() < 0.25 and len(live) < 9:
            # between tree phases every sequence is committed: fork one (at a random position or its end), or drop one
            parent = int(rng.choice(live))
            pos = int(rng.integers(1, length[parent] + 1)) if rng.random() < 0.6 else -1
            n = int(rng.integers(1, 12))
            p.forward([((next_id, parent, pos), n)])
            length[next_id] = (length[parent] if pos == -1 else pos) + n
            next_id += 1
            continue
        if forks and rng.random() < 0.08 and len(live) > 3:
            s_ = int(rng.choice(live))
            p.op(op="remove", seq=s_)
            del length[s_]
            continue
        k = int(rng.integers(1, len(live) + 1))
        sel = sorted(int(x) for x in rng.choice(live, size=k, replace=False))
        if rng.random() < 0.35:
            batch = [(s_, 1) for s_ in sel] if rng.random() < 0.6 else [(s_, int(rng.integers(1, 9))) for s_ in sel]
            p.forward(batch)
            for s_, n in batch:
                length[s_] += n
            continue
        rounds = int(rng.integers(1, 4))
        trees = {s_: [] for s_ in sel}
        for r in range(rounds):
            lens = []
            for s_ in sel:
                t = trees[s_]
                m = int(rng.integers(1, 7))
                for _j in range(m):
                    if not t:
                        t.append(-1)
                    elif r == 0 and rng.random() < 0.1:
                        t.append(-1)                               # a second root, as in the reference's own test
                    else:
                        t.append(int(rng.integers(0, len(t))))
                lens.append(m)
            leaves = None
            if r == rounds - 1:
                # CommitAcceptedTokenTreeNodes pops `last append length - (depth + 1)` (paged_kv_cache.cc:1646-1650) and looks
                # the path up in the LAST round's append_position_map: after several rounds only the first root (or
                # nothing) is a path it handles; a single-round tree commits any root-to-node path
                if rounds == 1:
                    leaves = [int(rng.integers(-1, len(trees[s_]))) for s_ in sel]
                else:
                    leaves = [int(rng.integers(-1, 1)) for s_ in sel]
            p.forward(list(zip(sel, lens)), trees=[trees[s_] for s_ in sel], leaves=leaves)
            last_lens = lens
        for s_, leaf, m_last in zip(sel, leaves, last_lens):
            depth, node = 0, leaf
            while node != -1:
                depth += 1
                node = trees[s_][node]
            length[s_] += len(trees[s_]) - (m_last - depth)
        if rng.random() < 0.3:
            p.op(op="query")
    p.dump_all({s_: min(n, 48) for s_, n in length.items()})
    p.op(op="query")
    return p


def prog_random_sliding(seed):
    """Randomised sliding-window program: sequences with random (window, sink) pairs, prefill chunks and many decode
    steps (so that every window slides over several pages and pages are recycled), forks inside the sink, removals."""
    rng = np.random.default_rng(seed)
    p = Program()
    length, window, sink = {}, {}, {}
    forked = set()
    next_id = 0

    def new_seq():
        nonlocal next_id
        sid = next_id
        next_id += 1
        w = int(rng.integers(8, 60))
        sk = int(rng.integers(0, w))
        p.op(op="add", seq=sid)
        p.op(op="enable_sw", seq=sid, window=w, sink=sk)
        length[sid], window[sid], sink[sid] = 0, w, sk
        return sid

    for _ in range(4):
        new_seq()
    for _ in range(45):
        live = sorted(length)
        r = rng.random()
        if r < 0.12 and len(live) < 8:
            sid = new_seq()
            n = int(rng.integers(1, 40))
            p.forward([(sid, n)])
            length[sid] += n
        elif r < 0.20 and len(live) < 8:
            # a sliding-window sequence may sit on at most two blocks: beyond that the reference merges the trailing blocks
            # into its last depth but keeps the LAST block's sink size / window offset for the merged page list
            # (paged_kv_cache.cc:884-1100), so its own plan reads slots nobody wrote.  Hence: fork a sequence once, and
            # never fork a child.
            cands = [s_ for s_ in live if 0 < length[s_] <= sink[s_] and s_ not in forked]
            if cands:
                parent = int(rng.choice(cands))
                child = next_id
                next_id += 1
                pos = int(rng.integers(1, length[parent] + 1)) if rng.random() < 0.5 else -1
                if (pos == -1 or pos == length[parent]) and length[parent] % 16 == 0:
                    # a fork at the page-aligned end gives the parent a new empty last block (paged_kv_cache.cc:635-645)
                    # without rebasing its last_block_attn_sink_size: the reference later trips its own
                    # ICHECK(block.seq_length >= block.sink_length) when that parent slides -- not a program it supports
                    pos = length[parent] - 1
                p.op(op="fork", parent=parent, child=child, pos=pos)
                forked.update((parent, child))
                base = length[parent] if pos == -1 else pos
                w = int(rng.integers(base + 2, base + 50))
                sk = int(rng.integers(base, w))
                p.op(op="enable_sw", seq=child, window=w, sink=sk)
                length[child], window[child], sink[child] = base, w, sk
                p.known.add(child)
                n = int(rng.integers(1, 20))
                p.forward([(child, n)])
                length[child] += n
        elif r < 0.85:
            k = int(rng.integers(1, len(live) + 1))
            sel = [int(x) for x in rng.choice(live, size=k, replace=False)]
            if rng.random() < 0.75:
                batch = [(s_, 1) for s_ in sel]
            else:
                batch = [(s_, int(rng.integers(1, 30))) for s_ in sel]
            p.forward(batch)
            for s_, n in batch:
                length[s_] += n
        elif len(live) > 2:
            s_ = int(rng.choice(live))
            p.op(op="remove", seq=s_)
            del length[s_], window[s_], sink[s_]
        if rng.random() < 0.15:
            p.op(op="query")
    for s_ in sorted(length):
        n = min(length[s_], window[s_])
        if n > 0:
            p.op(op="debug_get_kv", seq=s_, start=0, end=n)
    p.op(op="query")
    return p


def prog_shared_kv():
    """prefill, chunked prefill, decode and a fork, every step followed by a shared-KV query of the same layer."""
    p = Program()
    ln = {}
    for b in [[(0, 3)], [(0, 2)], [(0, 1)], [(1, 37), (2, 18)], [(0, 30), (1, 1), (2, 16)], [(0, 1), (1, 1), (2, 1)],
              [((3, 1, 20), 9)], [(0, 1), (3, 1)], [(1, 1), (2, 1), (3, 5)]]:
        p.forward(list(b), shared=True)
        for item, n in b:
            if isinstance(item, tuple):
                sid, _, pos = item
                ln[sid] = pos + n
            else:
                ln[item] = ln.get(item, 0) + n
    p.dump_all(ln)
    return p


def prog_self_cross_merge():
    """self_attention + cross_attention + merge_attn_output_inplace on an empty cache (nothing to cross-attend), on
    cached sequences, on a mixed batch (one fresh sequence), and across a fork (two block depths)."""
    p = Program()
    p.op(op="add", seq=0)
    p.op(op="add", seq=1)
    p.forward_split([(0, 5), (1, 19)])
    p.forward([(0, 21), (1, 40)])
    p.forward_split([(0, 7), (1, 2)])
    p.op(op="add", seq=2)
    p.forward_split([(0, 3), (2, 33), (1, 16)])
    p.forward([(0, 1), (1, 1)])
    p.forward([((3, 1, 25), 6), ((4, 1, -1), 18)])
    p.forward_split([(3, 4), (4, 9), (1, 2), (0, 17)])
    p.forward([(0, 1), (1, 1), (3, 1), (4, 1)])
    p.dump_all({0: 23, 1: 42, 3: 32, 4: 60})
    p.op(op="query")
    return p


SCENARIOS = {
    # name: (program builder, cache kwargs)
    "random_tree_a": (lambda: prog_random_tree(21), dict(rope_mode=1)),
    "random_tree_b": (lambda: prog_random_tree(22), dict(rope_mode=0, num_layers=2)),
    "random_sliding_a": (lambda: prog_random_sliding(31), dict(rope_mode=2, support_sliding_window=True)),
    "random_sliding_b": (lambda: prog_random_sliding(32), dict(rope_mode=1, support_sliding_window=True)),
    # a pipeline stage that owns layers [2, 4) of a 4-layer model (layer_indptr = [2, 4]): attention calls carry GLOBAL
    # layer ids, attn_kinds has one entry per model layer (layer 2 plain, layer 3 sliding)
    "layer_offset": (prog_shared_kv, dict(rope_mode=0, num_layers=2, layer_begin=2, attn_kinds=[3, 3, 0, 3],
                                          layer_sliding_window_size=32)),
    "layer_sliding": (prog_prefill_and_decode, dict(rope_mode=0, num_layers=2, attn_kinds=[3, 0], layer_sliding_window_size=20)),
    "layer_sliding_inline_rope": (prog_prefill_and_decode, dict(rope_mode=2, num_layers=2, attn_kinds=[0, 3],
                                                                layer_sliding_window_size=35)),
    "shared_kv": (prog_shared_kv, dict(rope_mode=2, num_layers=2)),
    # window 32 > the longest chunk appended behind cached tokens (30): a query farther than the window from EVERY cached
    # key makes the reference's CPU prefill kernel return NaN (0 / 0 on the fully masked row), which pins nothing
    "shared_kv_layer_sliding": (prog_shared_kv, dict(rope_mode=0, num_layers=2, attn_kinds=[3, 0],
                                                     layer_sliding_window_size=32)),
    "self_cross_merge": (prog_self_cross_merge, dict(rope_mode=0)),
    "self_cross_merge_inline_rope": (prog_self_cross_merge, dict(rope_mode=2, num_layers=2)),
    "random_a": (lambda: prog_random(11), dict(rope_mode=1)),
    "random_b": (lambda: prog_random(12), dict(rope_mode=0)),
    "random_c": (lambda: prog_random(13), dict(rope_mode=2, num_layers=2)),
    "prefill_and_decode": (prog_prefill_and_decode, dict(rope_mode=1)),
    "remove_and_popn": (prog_remove_and_popn, dict(rope_mode=1)),
    "fork": (prog_fork, dict(rope_mode=1)),
    "unlimited_depth": (prog_unlimited_depth, dict(rope_mode=0)),
    "prefill_and_decode_inline_rope_2layers": (prog_prefill_and_decode, dict(rope_mode=2, num_layers=2)),
    "sliding_window": (prog_sliding_window, dict(rope_mode=2, support_sliding_window=True)),
    "sliding_window_fork": (prog_sliding_window_fork, dict(rope_mode=2, support_sliding_window=True)),
    "tree_attn": (prog_tree_attn, dict(rope_mode=1)),
}
BASE = dict(num_layers=1, num_qo_heads=4, num_kv_heads=1, head_dim=128, dtype="float16", reserved_nseq=32,
            max_total_seq=2048, prefill_chunk=512, page_size=16, rope_scale=1.0, rope_theta=1e4,
            layer_sliding_window_size=None, attn_kinds=None, layer_begin=0)


def q2_for(seed, num_layers, n, hq, d, dtype="float16"):
    """The second (shared-KV) query of a step, [num_layers, n, hq, d]."""
    rng = np.random.default_rng(seed + 500000)
    return rng.random((num_layers, n, hq, d), dtype=np.float32).astype(dtype)


def run_reference(name):
    builder, kw = SCENARIOS[name]
    cfg = dict(BASE)
    cfg.update(kw)
    meta, arrays = capture(name, builder(), cfg)
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"kvcache_{name}.npz")
    np.savez_compressed(path, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(f"{name}: {len(meta['ops'])} ops, {os.path.getsize(path) / 1024:.0f} KiB")


def capture(name, prog, cfg, kernels=None):
    """Runs `prog` on the reference's cache; returns (meta, arrays) as stored in tests/golden/kvcache_<name>.npz.
    kernels: compiled callbacks of an earlier RefCache with the same configuration (RefCache.fns), to skip the rebuild."""
    import tvm
    import tvm_ffi
    from refenv import RefCache

    if cfg["attn_kinds"] is not None and any(k != 0 for k in cfg["attn_kinds"]):
        # DebugGetKV only takes all-MHA caches (paged_kv_cache.cc:1715-1717): the dump is replaced by the error check
        dumped = [o["seq"] for o in prog.ops if o["op"] == "debug_get_kv"]   # sequences alive at the end of the program
        prog.ops = [o for o in prog.ops if o["op"] != "debug_get_kv"]
        if dumped:
            prog.ops.append({"op": "debug_get_kv_rejected", "seq": dumped[0]})
    rc = RefCache(**cfg, kernels=kernels)
    L, hq, hkv, d = cfg["num_layers"], cfg["num_qo_heads"], cfg["num_kv_heads"], cfg["head_dim"]
    arrays = {}
    results = []
    for idx, op in enumerate(prog.ops):
        del rc.trace[:]
        res = {}
        k = op["op"]
        if k == "clear":
            rc.call("kv_state_clear")
        elif k == "add":
            rc.call("kv_state_add_sequence", op["seq"])
        elif k == "remove":
            rc.call("kv_state_remove_sequence", op["seq"])
        elif k == "fork":
            rc.call("kv_state_fork_sequence", op["parent"], op["child"], op["pos"])
        elif k == "popn":
            rc.call("kv_state_popn", op["seq"], op["n"])
        elif k == "enable_sw":
            rc.call("attention_kv_cache_enable_sliding_window_for_seq", op["seq"], op["window"], op["sink"])
        elif k == "commit":
            rc.call("attention_kv_cache_commit_accepted_token_tree_nodes", tvm_ffi.Shape(op["seq_ids"]),
                    tvm_ffi.Shape(op["leaves"]))
        elif k == "query":
            res["empty"] = bool(rc.call("attention_kv_cache_empty"))
            res["num_available_pages"] = int(rc.call("attention_kv_cache_get_num_available_pages"))
            res["total_sequence_length"] = int(rc.call("attention_kv_cache_get_total_sequence_length"))
        elif k == "debug_get_kv":
            n = op["end"] - op["start"]
            kk = tvm.runtime.empty((L, n, hkv, d), cfg["dtype"], device=rc.dev)
            vv = tvm.runtime.empty((L, n, hkv, d), cfg["dtype"], device=rc.dev)
            rc.call("attention_kv_cache_debug_get_kv", op["seq"], op["start"], op["end"], kk, vv)
            arrays[f"k_{idx}"] = kk.numpy()
            arrays[f"v_{idx}"] = vv.numpy()
        elif k == "debug_get_kv_rejected":
            kk = tvm.runtime.empty((L, 1, hkv, d), cfg["dtype"], device=rc.dev)
            try:
                rc.call("attention_kv_cache_debug_get_kv", op["seq"], 0, 1, kk, kk)
                raise AssertionError("the reference accepted DebugGetKV on a non-MHA layer")
            except tvm.error.InternalError as e:
                assert "Only MHA is supported for DebugGetKV" in str(e)
        elif k == "forward":
            tree = tvm_ffi.Shape(op["tree"]) if op["tree"] is not None else None
            rc.call("kv_state_begin_forward", tvm_ffi.Shape(op["seq_ids"]), tvm_ffi.Shape(op["lens"]), tree)
            n = sum(op["lens"])
            qkv = qkv_for(op["seed"], L, n, hq, hkv, d, cfg["dtype"])
            outs = []
            q2 = q2_for(op["seed"], L, n, hq, d, cfg["dtype"]) if op.get("shared") else None
            shared_outs = []
            for layer in range(L):
                o = tvm.runtime.empty((n, hq, d), cfg["dtype"], device=rc.dev)
                rc.call("attention_kv_cache_attention_with_fused_qkv", cfg["layer_begin"] + layer, d ** -0.5,
                        tvm.runtime.tensor(qkv[layer], device=rc.dev), o)
                outs.append(o.numpy())
                if q2 is not None:
                    # rope is none or inline in these scenarios, so the step's raw k / v are the "current" k / v
                    assert cfg["rope_mode"] != 1
                    o2 = tvm.runtime.empty((n, hq, d), cfg["dtype"], device=rc.dev)
                    rc.call("attention_kv_cache_attention_with_shared_kv", cfg["layer_begin"] + layer, d ** -0.5,
                            tvm.runtime.tensor(q2[layer], device=rc.dev),
                            tvm.runtime.tensor(np.ascontiguousarray(qkv[layer][:, hq:hq + hkv]), device=rc.dev),
                            tvm.runtime.tensor(np.ascontiguousarray(qkv[layer][:, hq + hkv:]), device=rc.dev), o2)
                    shared_outs.append(o2.numpy())
            rc.call("kv_state_end_forward")
            arrays[f"o_{idx}"] = np.stack(outs)
            if shared_outs:
                arrays[f"os_{idx}"] = np.stack(shared_outs)
            res["num_available_pages"] = int(rc.call("attention_kv_cache_get_num_available_pages"))
        elif k == "forward_split":
            rc.call("kv_state_begin_forward", tvm_ffi.Shape(op["seq_ids"]), tvm_ffi.Shape(op["lens"]), None)
            n = sum(op["lens"])
            qkv = qkv_for(op["seed"], L, n, hq, hkv, d, cfg["dtype"])
            merged, merged_lse, selfs, crosses = [], [], [], []
            for layer in range(L):
                dev = rc.dev
                q = tvm.runtime.tensor(np.ascontiguousarray(qkv[layer][:, :hq]), device=dev)
                kk = tvm.runtime.tensor(np.ascontiguousarray(qkv[layer][:, hq:hq + hkv]), device=dev)
                vv = tvm.runtime.tensor(np.ascontiguousarray(qkv[layer][:, hq + hkv:]), device=dev)
                o_self = tvm.runtime.tensor(np.zeros((n, hq, d), cfg["dtype"]), device=dev)
                lse_self = tvm.runtime.tensor(np.full((n, hq), -5e4, "float32"), device=dev)
                # cross_attention leaves (o, lse) untouched when no sequence of the batch has a cached page
                o_cross = tvm.runtime.tensor(np.zeros((n, hq, d), cfg["dtype"]), device=dev)
                lse_cross = tvm.runtime.tensor(np.full((n, hq), -5e4, "float32"), device=dev)
                rc.call("attention_kv_cache_self_attention", cfg["layer_begin"] + layer, d ** -0.5, q, kk, vv, o_self, lse_self)
                rc.call("attention_kv_cache_cross_attention", cfg["layer_begin"] + layer, d ** -0.5, q, o_cross, lse_cross)
                selfs.append(o_self.numpy())
                crosses.append(o_cross.numpy())
                ret = rc.call("attention_kv_cache_merge_attn_output_inplace", o_self, lse_self, o_cross, lse_cross)
                assert len(ret) == 2
                merged.append(o_self.numpy())
                merged_lse.append(lse_self.numpy())
            rc.call("kv_state_end_forward")
            arrays[f"o_{idx}"] = np.stack(merged)
            arrays[f"lse_{idx}"] = np.stack(merged_lse)
            arrays[f"oself_{idx}"] = np.stack(selfs)
            arrays[f"ocross_{idx}"] = np.stack(crosses)
            res["num_available_pages"] = int(rc.call("attention_kv_cache_get_num_available_pages"))
        else:
            raise ValueError(k)
        res["trace"] = [dict(r) for r in rc.trace]
        results.append(res)
    meta = {"name": name, "config": cfg, "ops": prog.ops, "results": results,
            "reference": "apache/tvm @ /root/reference, C++ PagedAttentionKVCacheObj + CPU TIR kernels (c target, gcc -O3)"}
    capture.last_kernels = rc.raw_fns
    return meta, arrays


if __name__ == "__main__":
    names = sys.argv[1:] or list(SCENARIOS)
    for nm in names:
        run_reference(nm)
