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
