"""Host-side model of the backward kernels' mbarrier / buffer protocols (tools/bwd_protocol_sim.py): random schedules of
the issuer, the compute warps, the drain warps and the reducer against an in-order tensor pipe and out-of-order TMA loads.
The shipped protocols must be clean under every schedule, including one warp that lags a whole step behind; the variants
that were wrong on the way must be caught:
  * dK/dV kernel with two S^T buffers (head_dim <= 64) but SHARED "P^T published" barriers -- a fast warp's arrival for step
    i+1 completes step i's phase without the slow warp, dV(i) reads a P^T that is not there (found by this model, fixed
    before it was seen on hardware);
  * the same with one S^T barrier for both buffers -- a slow warp is lapped and waits forever;
  * the fused backward in half steps without the issuer's wait for bar_dqfree -- dQ^T(h) overwrites dQ^T(h-1) before the
    drain warps have read it (seen on the GPU as a hang under ncu's replay passes)."""
import os
import re
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import bwd_protocol_sim as sim  # noqa: E402

SEEDS = range(6)


@pytest.mark.parametrize("two_s", [False, True])
@pytest.mark.parametrize("slow", [0, 300])
def test_dkv_kernel_protocol_is_clean(two_s, slow):
    for seed in SEEDS:
        ok, detail = sim.dkvt(two_s=two_s, seed=seed, slow_warp=slow)
        assert ok, detail


def test_dkv_kernel_shared_p_barriers_are_caught():
    assert not any(sim.dkvt(two_s=True, seed=s, per_buffer_pbar=False, slow_warp=300)[0] for s in SEEDS)


def test_dkv_kernel_shared_s_barrier_is_caught():
    assert not any(sim.dkvt(two_s=True, seed=s, per_buffer_bar=False, slow_warp=300)[0] for s in SEEDS)


@pytest.mark.parametrize("slow", [0, 300])
def test_dq_kernel_protocol_is_clean(slow):
    """dQ kernel v2: three issuer threads whose commits only cover their own MMAs."""
    for seed in SEEDS:
        ok, detail = sim.dq2(seed=seed, slow_warp=slow)
        assert ok, detail


@pytest.mark.parametrize("drain_delay,slow", [(0, 0), (400, 0), (0, 300)])
def test_fused2_protocol_is_clean(drain_delay, slow):
    for seed in SEEDS:
        ok, detail = sim.fused2(seed=seed, drain_delay=drain_delay, slow_warp=slow)
        assert ok, detail


def test_fused2_without_dqfree_wait_is_caught():
    assert not any(sim.fused2(wait_dqfree=False, seed=s, drain_delay=400)[0] for s in SEEDS)


def test_kernels_carry_the_rules_the_model_checks():
    """The model is hand-written beside the kernels: pin the lines it stands for."""
    bwd = open(os.path.join(ROOT, "aule-attention_b200", "csrc", "kernels", "attn_bwd_sm100.cu")).read()
    body = bwd[bwd.index("void bwd_dkv_t_body("):bwd.index("// dQ kernel: one CTA per")]
    assert re.search(r"mbar_arrive\(half \? \(odd_buf \? bar_pb1 : bar_pb\) : \(odd_buf \? bar_p1 : bar_p\)\)", body)
    assert "wait(odd_b ? bar_p1 : bar_p, p_par)" in body and "wait(odd_b ? bar_pb1 : bar_pb, p_par)" in body
    assert "mbar_wait(odd_buf ? bar_s1 : bar_s, two_s ? ((step >> 1) & 1) : (step & 1))" in body
    f2 = open(os.path.join(ROOT, "aule-attention_b200", "csrc", "kernels", "attn_bwd_fused2_sm100.cu")).read()
    assert "if (h > 0) wait(bar_dqfree, (h - 1) & 1);" in f2
