"""The arithmetic behind the multi-CTA partition of the SAH builder (candela_b200/csrc/builder.cu, split_rank_kernel /
split_gather_kernel), restated in Python and checked against the reference's sequential loop
(Source/Core/BVH/BVHConstructor.cpp:532-549: `if (centroid < border) swap(refs[i], refs[mid++])`).

The GPU tests compare whole buffers with the oracle; this file pins the closed form itself on flag strings chosen to
stress it (long runs, a single right reference in front, periodic patterns), and bounds the length of the dependent
chain the gather pass follows."""
import math

import numpy as np
import pytest


def lomuto(flags):
    """The reference's loop on the identity arrangement; returns (arrangement, mid)."""
    a = list(range(len(flags)))
    mid = 0
    for i in range(len(flags)):
        if flags[a[i]]:          # position i still holds its original element when the loop gets there
            a[i], a[mid] = a[mid], a[i]
            mid += 1
    return a, mid


def closed_form(flags, run_jumping=True):
    """split_rank + split_gather: left references land at their rank; position P >= nL keeps a right reference, and
    otherwise receives the right reference found by following q -> rank(q).  Returns (arrangement, nL, longest chain)."""
    n = len(flags)
    flags = np.asarray(flags, bool)
    rank = np.cumsum(flags) - flags               # exclusive count of left references
    n_left = int(flags.sum())
    rpos = np.flatnonzero(~flags)                 # rpos[j] = position of the j-th right reference
    out = [-1] * n
    for p in np.flatnonzero(flags):
        out[rank[p]] = int(p)
    longest = 0
    for P in range(n_left, n):
        p, steps = P, 0
        while flags[p]:
            c = p - rank[p]                       # right references before p
            assert c >= 1
            if run_jumping:
                run0 = rpos[c - 1] + 1            # first position of the run of left references around p
                p -= (p - run0) // c * c          # the hops that stay inside the run ...
            p -= c                                # ... and the one that leaves it
            steps += 1
        out[P] = int(p)
        longest = max(longest, steps)
    return out, n_left, longest


def patterns():
    rng = np.random.default_rng(7)
    yield "empty", []
    yield "all_left", [1] * 37
    yield "all_right", [0] * 37
    yield "one_right_in_front", [0] + [1] * 500
    yield "one_left_at_the_end", [0] * 500 + [1]
    yield "left_then_right", [1] * 100 + [0] * 100
    yield "right_then_left", [0] * 100 + [1] * 100
    yield "alternating", [i & 1 for i in range(301)]
    yield "period_4", [int(i % 4 != 0) for i in range(1000)]
    yield "rows", ([1] * 95 + [0] * 5) * 40            # a scanline grid cut near one edge
    yield "growing_runs", sum(([0] + [1] * k for k in range(1, 60)), [])
    for n in (1, 2, 3, 33, 257, 2049):
        for p_left in (0.02, 0.5, 0.98):
            yield f"random_{n}_{p_left}", list((rng.random(n) < p_left).astype(int))


@pytest.mark.parametrize("name,flags", list(patterns()), ids=[n for n, _ in patterns()])
def test_closed_form_equals_the_sequential_loop(name, flags):
    want, mid = lomuto(flags)
    for jumping in (False, True):
        got, n_left, _ = closed_form(flags, jumping)
        assert n_left == mid
        assert got == want, name
    assert sorted(want) == list(range(len(flags)))
    assert want[:mid] == [i for i, f in enumerate(flags) if f]   # Lomuto keeps the left references in order


def test_chain_length_with_and_without_run_jumping():
    """A single right reference in front of n left ones: the plain chain takes n hops (130 ms of dependent loads at
    262k references), the run-jumping one a single iteration.  The worst case left is ~sqrt(2 n) runs visited in a row."""
    n = 4000
    flags = [0] + [1] * n
    assert closed_form(flags, run_jumping=False)[2] == n
    assert closed_form(flags, run_jumping=True)[2] == 1
    growing = sum(([0] + [1] * k for k in range(1, 90)), [])          # run k is k long and has k right references before it
    _, _, longest = closed_form(growing, run_jumping=True)
    assert longest <= 2 * math.isqrt(2 * len(growing)) + 2
    rng = np.random.default_rng(3)
    for n in (1000, 20000):
        for p_left in (0.1, 0.5, 0.9, 0.99):
            f = list((rng.random(n) < p_left).astype(int))
            assert closed_form(f, True)[2] <= 2 * math.isqrt(2 * n) + 2


# ---------------------------------------------------------------------------------------------------------------------
# The bin-free split search of level_step_tiny_kernel (ranges of <= 9 references) against the sequential search of
# SearchSAHPlaneBinned (BVHConstructor.cpp:276-363), both restated in float32 numpy (no FMA, like -ffp-contract=off).
F = np.float32
BINS = 64
SENT_MAX, SENT_MIN, INF_COST = F(10000000.0), F(-10000000.0), F(1e29)


def _area(mn, mx):
    e = mx - mn                                            # Bounds::GetArea, BVHConstructor.h:41-44
    return F(F(e[0] * e[1]) + F(e[1] * e[2])) + F(e[2] * e[0])


def _bin_of(c, lo, scale):
    return min(BINS - 1, max(0, int(F(F(c - lo) * scale))))


def sequential_search(node_mn, node_mx, boxes_mn, boxes_mx, cents):
    best, axis, border = INF_COST, 0, node_mn[0]
    for ax in range(3):
        lo, hi = node_mn[ax], node_mx[ax]
        if lo == hi:
            continue
        extent = F(hi - lo)
        scale = F(F(BINS) / extent)
        cnt = np.zeros(BINS, int)
        bmn = np.full((BINS, 3), SENT_MAX, F)
        bmx = np.full((BINS, 3), SENT_MIN, F)
        for k in range(len(cents)):
            b = _bin_of(cents[k][ax], lo, scale)
            cnt[b] += 1
            bmn[b] = np.minimum(bmn[b], boxes_mn[k])
            bmx[b] = np.maximum(bmx[b], boxes_mx[k])
        step = F(extent / F(BINS))
        for i in range(BINS - 1):
            lmn, lmx = bmn[: i + 1].min(0), bmx[: i + 1].max(0)
            rmn, rmx = bmn[i + 1:].min(0), bmx[i + 1:].max(0)
            cost = F(F(F(int(cnt[: i + 1].sum())) * _area(lmn, lmx)) + F(F(int(cnt[i + 1:].sum())) * _area(rmn, rmx)))
            if cost < best:
                best, axis, border = cost, ax, F(lo + F(step * F(i + 1)))
    return axis, border


def candidate_search(node_mn, node_mx, boxes_mn, boxes_mx, cents):
    """One (axis, candidate) pair per lane: candidates are i = 0 and the references' own bins; minimum over (cost, axis, i)."""
    pairs = []
    n = len(cents)
    for ax in range(3):
        lo, hi = node_mn[ax], node_mx[ax]
        if lo == hi:
            continue
        extent = F(hi - lo)
        scale = F(F(BINS) / extent)
        bins = [_bin_of(cents[k][ax], lo, scale) for k in range(n)]
        for ci in [0] + bins:
            if ci >= BINS - 1:
                continue
            left = [k for k in range(n) if bins[k] <= ci]
            right = [k for k in range(n) if bins[k] > ci]

            def side(idx):
                if not idx:
                    return 0, np.full(3, SENT_MAX, F), np.full(3, SENT_MIN, F)
                return len(idx), boxes_mn[idx].min(0), boxes_mx[idx].max(0)
            (nl, lmn, lmx), (nr, rmn, rmx) = side(left), side(right)
            cost = F(F(F(nl) * _area(lmn, lmx)) + F(F(nr) * _area(rmn, rmx)))
            if cost == cost:
                pairs.append((cost, ax, ci, F(lo + F(F(extent / F(BINS)) * F(ci + 1)))))
    if not pairs:
        return 0, node_mn[0]
    cost, ax, ci, border = min(pairs, key=lambda p: (p[0], p[1], p[2]))
    return (ax, border) if cost < INF_COST else (0, node_mn[0])


def _random_range(rng, n, kind):
    c = rng.uniform(-4, 4, size=(n, 3)).astype(F)
    if kind == "duplicates":
        c[:] = c[0]
    elif kind == "flat":
        c[:, 1] = F(0.5)                                   # all boxes in one plane: that axis may be degenerate
    elif kind == "lattice":
        c = np.round(c).astype(F)                          # many equal centroids and equal costs
    half = rng.uniform(0, 0.6, size=(n, 3)).astype(F) if kind != "flat" else np.abs(rng.normal(0, 0.3, (n, 3))).astype(F) * np.array([1, 0, 1], F)
    mn, mx = (c - half).astype(F), (c + half).astype(F)
    cents = ((mn + mx) / F(2)).astype(F)                   # Bounds::GetCenter
    return mn.min(0), mx.max(0), mn, mx, cents


@pytest.mark.parametrize("kind", ["random", "duplicates", "flat", "lattice"])
def test_candidate_split_search_equals_the_sequential_search(kind):
    rng = np.random.default_rng(11)
    for trial in range(150):
        n = int(rng.integers(3, 10))
        node_mn, node_mx, mn, mx, cents = _random_range(rng, n, kind)
        want = sequential_search(node_mn, node_mx, mn, mx, cents)
        got = candidate_search(node_mn, node_mx, mn, mx, cents)
        assert got[0] == want[0] and got[1].tobytes() == want[1].tobytes(), (kind, trial, got, want)


def test_closed_form_on_arbitrary_flag_strings():
    """Property check (hypothesis): any flag string, with and without run jumping."""
    hypothesis = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.booleans(), max_size=300))
    def check(flags):
        want, mid = lomuto(flags)
        for jumping in (False, True):
            got, n_left, _ = closed_form(flags, jumping)
            assert (got, n_left) == (want, mid)

    check()
