"""Executable specification (CPU, Python integers) of the two transformations the fused step kernels apply to the
oracle's systematic-resampling formula (stochy.jl_b200/csrc/sb_device.cuh, sb_pull_tile / sb_pull_ancestors):

 1. the nested-floor form: inside a thread the CDF is evaluated once in 64 bits, K = (baseb + (c R >> rb)) >> s, and the
    child index at the prefix c + a is K + ((Z0 + a R) >> (rb + s)) with Z0 = (D0 mod 2^s) 2^rb + (c R mod 2^rb);
 2. the boundary table (a warp skips its 256 parents when the child index at its prefix equals the one at the next
    warp's) + head markers + boundary owners + warp-local fill: every parent that owns children writes its index at its first
    child's slot; the owner of the first 256-slot boundary above a thread's first child writes one there too (all of
    them when a parent covers several boundaries); each warp of 256 child slots then fills from its own markers only.

Both must reproduce, slot for slot, the literal definition: child k belongs to the parent i with F(C_i) <= k < F(C_{i+1}),
F(C) = number of thresholds U + k 2^s strictly below C (strict `<` as in src/rand.jl:9)."""
import random

import pytest

TILE, ITEMS, WARP = 2048, 8, 32
SEEDS = {"uniformish": 11, "sparse": 12, "heavy": 13, "twolevel": 14}


def clamp(k, cw):
    return 0 if k <= 0 else (cw if k >= cw else k)


def literal_k(baseb, acc, rb, s, cw):
    return clamp((baseb + (acc >> rb)) >> s, cw)             # Python's >> on negative ints is the arithmetic shift


def thread_his(baseb, c, w, R, rb, s, cw):
    """hi[j] of one thread by the nested-floor form (what the kernel computes when 32 <= rb + s <= 62)."""
    acc = c * R
    D0 = baseb + (acc >> rb)
    K = min(D0 >> s, 1 << 29)
    Z = ((D0 & ((1 << s) - 1)) << rb) | (acc & ((1 << rb) - 1))
    lo0 = clamp(K, cw)
    his = []
    for wj in w:
        Z += wj * R
        assert Z < 1 << 64
        his.append(clamp(K + (Z >> (rb + s)), cw))
    return lo0, his


def random_case(rng, kind):
    s = rng.choice([32, 33, 36])
    rb = rng.randint(0, 62 - s)
    R = rng.randint(1 << 25, (1 << 26) - 1)
    scale = rng.choice([1, 1, 1, 64, 4096])
    if kind == "uniformish":
        w = [rng.randint(1, 1 << 27) // scale for _ in range(2 * TILE)]
    elif kind == "sparse":
        w = [(rng.randint(1, 1 << 27) if rng.random() < 0.03 else 0) for _ in range(2 * TILE)]
    elif kind == "heavy":
        w = [rng.randint(0, 1 << 12) for _ in range(2 * TILE)]
        for _ in range(rng.randint(1, 6)):
            w[rng.randrange(2 * TILE)] = (1 << 27) - 1
    else:                                                       # two levels, as the sticky HMM produces
        w = [rng.choice([(1 << 27) - 5, (1 << 22) + 3]) for _ in range(2 * TILE)]
    return s, rb, R, w


@pytest.mark.parametrize("kind", ["uniformish", "sparse", "heavy", "twolevel"])
def test_nested_floor_form_equals_literal_formula(kind):
    rng = random.Random(SEEDS[kind])
    for _ in range(300):
        s, rb, R, w = random_case(rng, kind)
        cw = rng.choice([TILE, TILE, 1500, 7])
        c = rng.randint(0, 1 << 38)
        baseb = rng.randint(-(1 << 50), 1 << 50)
        ws = w[:ITEMS]
        lo0, his = thread_his(baseb, c, ws, R, rb, s, cw)
        assert lo0 == literal_k(baseb, c * R, rb, s, cw)
        a = 0
        for j in range(ITEMS):
            a += ws[j]
            assert his[j] == literal_k(baseb, (c + a) * R, rb, s, cw)


def kernel_model(w, R, rb, s, U, c0, cw):
    """Ancestors of the children [c0, c0 + cw) from two staged parent tiles, the way the kernel computes them."""
    marker = [-1] * TILE
    for which in range(2):
        wt = w[which * TILE:(which + 1) * TILE]
        Ptb = (sum(w[:TILE]) * R >> rb) if which else 0       # rescaled prefix of the tile (same exponent: one shift)
        baseb = Ptb - U + (1 << s) - 1 - (c0 << s)
        # the boundary table (sb_pull_seg): child index at the 8 warp prefixes and at the tile's end; a warp whose two
        # entries coincide skips its 256 parents
        wpre = [sum(wt[:256 * g]) for g in range(8)]
        segk = [literal_k(baseb, wpre[g] * R, rb, s, cw) for g in range(8)] + [literal_k(baseb, sum(wt) * R, rb, s, cw)]
        c = 0
        for t in range(TILE // ITEMS):
            ws = wt[t * ITEMS:(t + 1) * ITEMS]
            base = which * TILE + t * ITEMS
            g = t // WARP
            if segk[g + 1] > segk[g] and sum(ws):
                lo0, his = thread_his(baseb, c, ws, R, rb, s, cw)
                if his[-1] > lo0:
                    lo = lo0
                    for j in range(ITEMS):
                        if lo < his[j]:
                            marker[lo] = base + j
                        lo = his[j]
                    bnd = (lo0 | 255) + 1
                    if his[-1] > bnd:
                        marker[bnd] = base + sum(1 for j in range(ITEMS - 1) if his[j] <= bnd)
                        if his[-1] > bnd + 256:
                            lo = lo0
                            for j in range(ITEMS):
                                q = (lo | 255) + 1
                                while q < his[j]:
                                    marker[q] = base + j
                                    q += 256
                                lo = his[j]
            c += sum(ws)
    anc = []
    for wp in range(TILE // 256):                               # one warp of children: its own 256 slots only
        prev = 0
        for slot in range(wp * 256, (wp + 1) * 256):
            prev = max(prev, marker[slot])
            anc.append(prev)
    return anc[:cw]


def literal_model(w, R, rb, s, U, c0, cw):
    anc = [0] * cw
    pre = 0
    Q_A = sum(w[:TILE])
    for i, wi in enumerate(w):
        which = i // TILE
        Ptb = (Q_A * R >> rb) if which else 0
        loc = pre - (Q_A if which else 0)
        baseb = Ptb - U + (1 << s) - 1 - (c0 << s)
        lo = literal_k(baseb, loc * R, rb, s, cw)
        hi = literal_k(baseb, (loc + wi) * R, rb, s, cw)
        for k in range(lo, hi):
            anc[k] = i
        pre += wi
    return anc


@pytest.mark.parametrize("kind", ["uniformish", "sparse", "heavy", "twolevel"])
def test_markers_and_warp_local_fill_equal_literal_ancestors(kind):
    rng = random.Random(1000 + SEEDS[kind])
    done = 0
    for _ in range(40):
        s, rb, R, w = random_case(rng, kind)
        total = sum(w) * R >> rb
        if total >> s < 64:
            continue                                            # fewer than 64 children over both tiles: nothing to pull
        # a children tile somewhere inside the range these two parent tiles feed, thresholds at U + k 2^s
        U = rng.randint(0, (1 << s) - 1)
        nchildren = (total - U - 1 >> s) + 1 if total > U else 0
        cw = min(TILE, nchildren)
        c0 = rng.randint(0, max(0, nchildren - cw))
        # children outside [first child of A, last child of B] belong to other tiles: restrict to those fed here
        got = kernel_model(w, R, rb, s, U, c0, cw)
        ref = literal_model(w, R, rb, s, U, c0, cw)
        assert got == ref
        done += 1
    assert done >= 10


# ---------------------------------------------------------------------------------------------------------------
# The speculative exact total (sb_spec_term, k_step_hmm / k_step_lg -> k_tilescan) and its sharded form (the GPUs'
# summary records, sb_signal_done): a model of the integer arithmetic.  The scan's first pass is
#     total = sum_b (Q_b << LS) >> (E - e_b),   E = max_b e_b over tiles with Q_b != 0,   LS = 24 - ceil_log2(tiles);
# the step kernels accumulate the same terms at the exponent E_top they expect, CTA by CTA and GPU by GPU, in any order.
E_EMPTY = -(1 << 30)


def scan_total(Q, e, LS):
    live = [eb for q, eb in zip(Q, e) if q]
    if not live:
        return E_EMPTY, 0
    E = max(live)
    return E, sum(((q << LS) >> (E - eb)) if E - eb < 63 else 0 for q, eb in zip(Q, e) if q)


def spec_term(q, eb, E_top, LS):
    if q == 0 or eb == E_EMPTY:
        return 0
    sh = E_top - eb
    return 0 if sh >= 63 else (q << LS) >> sh


@pytest.mark.parametrize("world", [1, 2, 8])
def test_speculative_total_equals_scan_total_or_is_refused(world):
    rng = random.Random(20 + world)
    for case in range(300):
        tiles_per_gpu = rng.choice([1, 2, 8, 64])
        nt = tiles_per_gpu * world
        LS = 24 - (0 if nt <= 1 else (nt - 1).bit_length())
        E_top = rng.randint(-40, 40)
        hit = rng.random() < 0.7                      # some tile reaches the exponent the step expected
        Q, e = [], []
        for b in range(nt):
            if rng.random() < 0.1:
                Q.append(0); e.append(E_EMPTY)        # a tile of zero weights
            else:
                Q.append(rng.randint(1, 2048 << 27))
                e.append(E_top - rng.choice([0, 0, 0, 1, 2, 7, 70]) - (0 if hit else 1))
        if hit and not any(q and eb == E_top for q, eb in zip(Q, e)):
            Q[0], e[0] = 12345, E_top
        E, total = scan_total(Q, e, LS)
        # per GPU: CTAs take the tiles round-robin and add their terms in any order; one record per GPU
        records = []
        for g in range(world):
            mine = list(range(g * tiles_per_gpu, (g + 1) * tiles_per_gpu))
            rng.shuffle(mine)
            live = [e[b] for b in mine if Q[b]]
            records.append((max(live) if live else E_EMPTY, E_top, sum(spec_term(Q[b], e[b], E_top, LS) for b in mine)))
        E_rec = max(r[0] for r in records)
        assert E_rec == E
        take = all(r[1] == records[0][1] for r in records) and records[0][1] == E_rec and E_rec != E_EMPTY
        assert take == (hit and E != E_EMPTY) or not any(Q)
        if take:
            assert sum(r[2] for r in records) == total
            assert total < 1 << 62                       # the scan's head-room (LS) holds for the speculative sum too
