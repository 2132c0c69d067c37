"""The balanced split-KV partition of decode.cu, restated in Python and checked for consistency on random batches (no GPU).

Two pieces of device code must agree without ever talking to each other:
  * decode_kernel: CTA k walks the page-heads [k * quota, (k + 1) * quota) of the line (sequence-major, then kv head, then
    page), cutting (sequence, head) segments only at its own boundaries, and writes a partial for every cut piece into slot
    2 k + (the piece continues into CTA k + 1 ? 1 : 0); a segment that lies inside one CTA is written directly;
  * decode_merge_kernel / seg_pieces: for (b, h) the pieces live in CTAs k0 .. k1 computed from page_indptr alone, piece c
    in slot 2 (k0 + c) + (c < nc - 1).
This model runs both on the same random inputs (empty sequences, one-page sequences, sequences spanning many CTAs, any
quota) and requires: every page-head is covered exactly once, slots are unique, and the merge side finds exactly the
partials the walk wrote -- the GPU tests then only have to show that the CUDA code implements these few lines."""
import numpy as np
import pytest


def kernel_walk(page_indptr, hkv, quota):
    """-> (direct: set of (b, h), partials: dict slot -> (b, h, pg0, pg1), covered: list of (b, h, page))"""
    B = len(page_indptr) - 1
    total = int(page_indptr[B]) * hkv
    grid = max(1, -(-total // quota))
    direct, partials, covered = set(), {}, []
    for k in range(grid):
        lin, lin_end = k * quota, min((k + 1) * quota, total)
        if lin >= lin_end:
            continue
        # the last sequence b with page_indptr[b] * hkv <= lin (binary search in the kernel)
        b = int(np.searchsorted(np.asarray(page_indptr[:B]) * hkv, lin, side="right") - 1)
        np_b = page_indptr[b + 1] - page_indptr[b]
        off = lin - page_indptr[b] * hkv
        h, pg0 = off // np_b, off % np_b
        while lin < lin_end:
            np_b = page_indptr[b + 1] - page_indptr[b]
            pg1 = np_b if np_b - pg0 <= lin_end - lin else pg0 + (lin_end - lin)
            covered += [(b, h, pg) for pg in range(pg0, pg1)]
            if pg0 == 0 and pg1 == np_b:
                assert (b, h) not in direct
                direct.add((b, h))
            else:
                slot = 2 * k + (1 if pg1 < np_b else 0)
                assert slot not in partials, "a CTA wrote one of its two slots twice"
                partials[slot] = (b, h, pg0, pg1)
            lin += pg1 - pg0
            pg0 = 0
            h += 1
            if h == hkv:
                h = 0
                b += 1
                while b < B and page_indptr[b + 1] == page_indptr[b]:
                    b += 1
    return direct, partials, covered


def merge_side(page_indptr, b, h, hkv, quota):
    """seg_pieces: -> list of slots (empty list: nothing to merge)"""
    p0, np_b = page_indptr[b], page_indptr[b + 1] - page_indptr[b]
    if np_b == 0:
        return []
    s0 = p0 * hkv + h * np_b
    k0 = s0 // quota
    nc = (s0 + np_b - 1) // quota - k0 + 1
    return [] if nc <= 1 else [2 * (k0 + c) + (1 if c < nc - 1 else 0) for c in range(nc)]


@pytest.mark.parametrize("seed", range(40))
def test_walk_and_merge_agree(seed):
    rng = np.random.default_rng(seed)
    B = int(rng.integers(1, 30))
    hkv = int(rng.choice([1, 2, 4, 8]))
    pages = [0 if rng.random() < 0.15 else int(rng.choice([1, 2, 3, int(rng.integers(1, 40)), int(rng.integers(40, 600))]))
             for _ in range(B)]
    indptr = [0]
    for n in pages:
        indptr.append(indptr[-1] + n)
    total = indptr[-1] * hkv
    quota = int(rng.choice([4, 8, 12, 16, 40, 444, max(4, (-(-total // 296) + 3) // 4 * 4)]))
    direct, partials, covered = kernel_walk(indptr, hkv, quota)
    want = [(b, h, pg) for b in range(B) for h in range(hkv) for pg in range(pages[b])]
    assert sorted(covered) == want, "every page-head exactly once"
    for b in range(B):
        for h in range(hkv):
            slots = merge_side(indptr, b, h, hkv, quota)
            if pages[b] == 0:
                assert slots == [] and (b, h) not in direct      # CTA 0 writes the empty result
            elif not slots:
                assert (b, h) in direct
            else:
                assert (b, h) not in direct
                pieces = [partials[s] for s in slots]            # KeyError = the merge would read a slot nobody wrote
                assert all(p[0] == b and p[1] == h for p in pieces)
                assert pieces[0][2] == 0 and pieces[-1][3] == pages[b]
                assert all(pieces[i][3] == pieces[i + 1][2] for i in range(len(pieces) - 1)), "pieces are contiguous, in order"
    used = {s for b in range(B) for h in range(hkv) for s in merge_side(indptr, b, h, hkv, quota)}
    assert used == set(partials), "no partial is left unread"


def test_host_plan_quota_and_grid():
    """decode_entry's host plan: one quota per resident CTA (2 per SM), never below 8 page-heads, a multiple of the 4 warps"""
    def plan(nnz, hkv, grid_max=296):
        total = nnz * hkv
        quota = max(-(-total // grid_max), 8)
        quota = -(-quota // 4) * 4
        return quota, max(1, -(-total // quota))

    assert plan(64 * 256, 8) == (444, 296)         # C2
    assert plan(256 * 512, 1) == (444, 296)        # C4 at tp 8
    assert plan(32 * 2048, 8) == (1772, 296)       # C5 decode
    assert plan(4 * 3, 8) == (8, 12)               # a tiny batch: fewer CTAs, not one-page pieces
    assert plan(0, 8) == (8, 1)                    # nothing cached: one CTA still writes the empty results
