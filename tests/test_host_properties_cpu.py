"""CPU: property tests (hypothesis) of the host-side logic against the oracle restatements -- visible-shape choice,
band grouping, window origins, cosine schedule."""
import math
import random

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from hsimae_b200 import feed, gwpca, host
from hsimae_b200.optim import CosineLRScheduler
from oracle import gwpca_oracle as G, hsimae_oracle as O


@settings(max_examples=150, deadline=None, derandomize=True)
@given(T=st.integers(2, 8), L=st.integers(2, 16), ratio=st.floats(0.0, 0.95), seed=st.integers(0, 1000))
def test_visible_shape_properties(T, L, ratio, seed):
    random.seed(seed)
    lt, ll = host.choose_visible_shape(T, L, ratio)
    state = random.getstate()
    random.seed(seed)
    assert (lt, ll) == O.choose_visible_shape(T, L, ratio)
    assert random.getstate() == state                         # same RNG consumption as the restated reference
    assert 2 <= lt <= T and 2 <= ll <= L
    target = (1 - ratio) * T * L
    best = min(abs(np.float32(target) - np.float32(t * l)) for t in range(2, T + 1) for l in range(2, L + 1))
    assert abs(np.float32(target) - np.float32(lt * ll)) <= best * (1 + 1e-6) + 1e-6


@settings(max_examples=200, deadline=None, derandomize=True)
@given(c=st.integers(8, 400), group=st.sampled_from([2, 4, 6, 8]))
def test_band_groups_partition_the_bands(c, group):
    g = gwpca.band_groups(c, group)
    assert g == G.band_groups(c, group)
    assert len(g) == 2 ** (group // 2)
    assert g[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(g, g[1:])) and g[-1][0] + g[-1][1] == c
    assert max(w for _, w in g) - min(w for _, w in g) <= group // 2      # halving keeps the widths within one per round


@settings(max_examples=200, deadline=None, derandomize=True)
@given(length=st.integers(9, 600), stride=st.sampled_from([1, 3, 9]))
def test_window_origins_cover_the_axis(length, stride):
    seq = feed.initial_seq(length, 9, stride)
    assert seq[0] == 0 or len(seq) == 1
    assert seq[-1] == length - 9 and np.all(seq >= 0) and np.all(seq + 9 <= length)
    covered = np.zeros(length, dtype=bool)
    for o in seq:
        covered[o:o + 9] = True
    assert covered.all()                                                    # every pixel row/column is inside some window
    assert np.all(np.diff(seq[:-1]) == 9 // stride)


@settings(max_examples=60, deadline=None, derandomize=True)
@given(iters=st.integers(20, 5000), base=st.floats(1e-5, 1e-1))
def test_cosine_schedule_properties(iters, base):
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=base)
    warm = int(math.ceil(iters * 0.05))
    s = CosineLRScheduler(opt, t_initial=iters, lr_min=1e-6, warmup_t=warm)
    assert opt.param_groups[0]["lr"] == 0.0                                  # starts at warmup_lr_init
    lrs = [s.get_lr(t)[0] for t in range(iters + 3)]
    assert all(b >= a for a, b in zip(lrs[:warm], lrs[1:warm]))             # linear warm-up
    assert all(b <= a + 1e-15 for a, b in zip(lrs[warm:], lrs[warm + 1:]))  # then monotone decay
    assert max(lrs) <= base * (1 + 1e-12) and abs(lrs[-1] - 1e-6) < 1e-12
