"""The stream-K work split of the decode GEMM, restated in Python and checked for the properties the kernel relies on.

Mirrors `Sched<true>` and the owner's contributor range in `swap_epilogue_tile` (isca-2025-lia_b200/csrc/gemm_shared.cuh):
CTA c of G streams the flat (tile, k-block) range [total*c/G, total*(c+1)/G); the CTA whose piece starts a tile (kb0 == 0)
owns it and adds the pieces of CTAs c+1 .. last_c in that fixed order, where last_c is computed in closed form
(one division) instead of by walking the CTAs.  No GPU, no library: this is host-checkable arithmetic.
"""
import random

import pytest


def spans(tiles_a, k_blocks, ncta):
    total = tiles_a * k_blocks
    return [(total * c // ncta, total * (c + 1) // ncta) for c in range(ncta)]


def work_items(tiles_a, k_blocks, ncta, cta):
    """Sched<true>::next for one CTA: (ta, kb0, kb1) pieces in order."""
    pos, end = spans(tiles_a, k_blocks, ncta)[cta]
    out = []
    while pos < end:
        ta = pos // k_blocks
        kb0 = pos - ta * k_blocks
        left = end - pos
        kb1 = kb0 + left if left < k_blocks - kb0 else k_blocks
        out.append((ta, kb0, kb1))
        pos += kb1 - kb0
    return out


def last_contributor_loop(tiles_a, k_blocks, ncta, cta, ta):
    total, tile_end = tiles_a * k_blocks, (ta + 1) * k_blocks
    last_c = cta
    while last_c + 1 < ncta and total * (last_c + 1) // ncta < tile_end:
        last_c += 1
    return last_c


def last_contributor_closed_form(tiles_a, k_blocks, ncta, cta, ta):
    total, tile_end = tiles_a * k_blocks, (ta + 1) * k_blocks
    return max(min(ncta - 1, (tile_end * ncta - 1) // total), cta)


SHAPES = [
    # tiles_a, k_blocks, CTAs: the decode projections of OPT-30B at TP1 / TP8 and the lm_head, on 148 SMs
    (168, 112, 148), (56, 112, 148), (224, 112, 148), (56, 448, 148), (393, 112, 148),
    (21, 112, 148), (56, 14, 148), (28, 112, 148), (56, 56, 148),
    (1, 4, 1), (1, 600, 148), (3, 5, 2), (7, 9, 5),
]


@pytest.mark.parametrize("tiles_a,k_blocks,ncta", SHAPES)
def test_every_k_block_of_every_tile_is_streamed_exactly_once(tiles_a, k_blocks, ncta):
    seen = [[0] * k_blocks for _ in range(tiles_a)]
    for cta in range(ncta):
        for ta, kb0, kb1 in work_items(tiles_a, k_blocks, ncta, cta):
            assert 0 <= kb0 < kb1 <= k_blocks
            for kb in range(kb0, kb1):
                seen[ta][kb] += 1
    assert all(v == 1 for row in seen for v in row)


@pytest.mark.parametrize("tiles_a,k_blocks,ncta", SHAPES)
def test_owner_sums_exactly_the_other_pieces_of_its_tile_in_k_order(tiles_a, k_blocks, ncta):
    pieces = {}                       # tile -> [(kb0, cta)]
    for cta in range(ncta):
        for ta, kb0, kb1 in work_items(tiles_a, k_blocks, ncta, cta):
            pieces.setdefault(ta, []).append((kb0, cta))
    for ta, lst in pieces.items():
        lst.sort()
        owner = lst[0][1]
        assert lst[0][0] == 0
        others = [c for _, c in lst[1:]]
        last_c = last_contributor_closed_form(tiles_a, k_blocks, ncta, owner, ta)
        assert others == list(range(owner + 1, last_c + 1)), (ta, owner, others, last_c)


def test_closed_form_matches_the_walk_on_random_shapes():
    rng = random.Random(7)
    for _ in range(20000):
        ncta = rng.randint(1, 160)
        tiles_a, k_blocks = rng.randint(1, 400), rng.randint(1, 500)
        cta = rng.randrange(ncta)
        for ta, kb0, _ in work_items(tiles_a, k_blocks, ncta, cta):
            if kb0 == 0:
                assert (last_contributor_closed_form(tiles_a, k_blocks, ncta, cta, ta)
                        == last_contributor_loop(tiles_a, k_blocks, ncta, cta, ta))


# ---- two-shot decode exchange: which rank reduces which row group (gemm_shared.cuh, swap_epilogue_tile / tp_two_shot_rows)
def owners_incremental(ta, world, iters):
    """The kernel's walk: one modulo per tile, then +1 with wrap per row group."""
    own, out = ta % world, []
    for _ in range(iters):
        out.append(own)
        own = 0 if own + 1 == world else own + 1
    return out


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("iters", [2, 4, 8, 16])
def test_row_group_owner_walk_equals_the_modulo_it_replaced(world, iters):
    for ta in range(0, 200):
        want = [(it + ta) % world for it in range(iters)]
        assert owners_incremental(ta, world, iters) == want
        for rank in range(world):
            mine = 0
            for it, o in enumerate(owners_incremental(ta, world, iters)):
                if o == rank:
                    mine |= 1 << it
            assert [bool((mine >> it) & 1) for it in range(iters)] == [w == rank for w in want]
        # every row group has exactly one reducing rank, and each rank reduces its share to within one group
        counts = [want.count(r) for r in range(world)]
        assert sum(counts) == iters and max(counts) - min(counts) <= 1


# ---- launch bookkeeping of the decode attention (ops.attn_decode_launches restates csrc/attn_decode.cu's split choice)
def test_decode_attention_launch_count_follows_the_split_choice(monkeypatch):
    import lia_b200  # noqa: F401
    from lia_b200 import ops
    monkeypatch.delenv("LIA_ATTN_CTAS_PER_SM", raising=False)
    n = lambda B, H, T, splits, ws=True: ops.attn_decode_launches(B, H, T, splits, ws, 148)  # noqa: E731
    assert n(64, 56, 288, 0) == 1          # headline, unsharded: one CTA per (b, h), the reference's rounding of p
    assert n(64, 28, 288, -6) == 1         # TP2
    assert n(64, 14, 288, -6) == 1         # TP4
    assert n(64, 7, 288, 0) == 1           # TP8 shapes under the unsharded default
    assert n(64, 7, 288, -6) == 2          # TP8 as the tensor-parallel model asks: two key ranges + the combine kernel
    assert n(64, 7, 200, -6) == 1          # fewer than 256 keys: never split below 128 keys per CTA
    assert n(8, 32, 300, 0) == 2           # small batch
    assert n(8, 32, 300, 0, ws=False) == 1  # no workspace: the fewest splits that fit
    assert n(64, 56, 288, 1) == 1 and n(64, 56, 288, 4) == 2
    monkeypatch.setenv("LIA_ATTN_CTAS_PER_SM", "6")
    assert n(64, 7, 288, 0) == 2
