"""Host-side model of the paged decode kernel's K/V ring (tools/paged_ring_sim.py): the stage count must be a
multiple of the consumer-warp count or a consumer can pass its parity wait on a stage whose previous fill is still in
flight (found on the GPU with NS = 10, then reproduced here).  The shipped configuration must satisfy the rule."""
import os
import re
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import paged_ring_sim as sim  # noqa: E402

ITEMS = [37] * 5 + [31]


@pytest.mark.parametrize("ns", [8, 12, 16, 24])
def test_ring_is_safe_when_stages_are_a_multiple_of_the_consumer_warps(ns):
    for seed in range(5):
        ok, detail = sim.simulate(ns, ITEMS, max_delay=60, seed=seed)
        assert ok, detail


def test_ring_with_ten_stages_reads_stale_tiles():
    bad = [sim.simulate(10, ITEMS, max_delay=60, seed=seed)[0] for seed in range(5)]
    assert not all(bad)


def test_shipped_configuration_obeys_the_rule():
    src = open(os.path.join(ROOT, "aule-attention_b200", "csrc", "kernels", "kernel_params.h")).read()
    paged = src[src.index("struct PagedCfg"):]
    ns128, ns64 = map(int, re.search(r"NS = \(D == 128\) \? (\d+) : (\d+);", paged).groups())
    nw = int(re.search(r"CONSUMERS = (\d+);", paged).group(1))
    assert ns128 % nw == 0 and ns64 % nw == 0
    assert "static_assert(NS % CONSUMERS == 0" in paged
