"""
CPU ORACLE for the fused-attention hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
path (``aule-attention_b200/``) never does: it fails loudly when the CUDA
library is missing.

Parity pinning: this restatement is checked against
  * outputs of the UNMODIFIED reference NumPy path
    (``/root/reference/python/aule/__init__.py:247-271``) generated in the
    build container by ``tests/golden/gen_golden.py`` and committed under
    ``tests/golden/``;
  * every known-answer vector the reference's own tests hold for this path
    (``src/attention_ref.zig:250-298``, ``tests/test_attention.zig:158-270``).
See ``tests/test_oracle.py``.

Functions
---------
cpu_attention            bit-for-bit restatement of the reference NumPy path
                         (same einsum / where / exp calls, fp32, scale ignored,
                         MHA only) -- python/aule/__init__.py:247-271
attention_ref            extended oracle: GQA broadcast, explicit scale, LSE,
                         selectable accumulation dtype (fp64 for budgeting)
                         -- semantics of python/aule/triton_flash.py:95-96,
                         :183-235 and src/attention_ref.zig:97-171
attention_rows           row-block evaluation so configs C/D can be checked on
                         sampled (b, h, rows) without materialising S x S
attention_bwd_ref        analytic dQ/dK/dV  -- python/aule/triton_flash.py:321-347
"""
import math

import numpy as np


# --------------------------------------------------------------------------
# 1. The reference NumPy path, restated call for call.
#    /root/reference/python/aule/__init__.py:247-271
# --------------------------------------------------------------------------
def cpu_attention(q, k, v, causal=True):
    """softmax(Q K^T / sqrt(D) + causal(-1e9)) V, materialised, input dtype.

    Follows python/aule/__init__.py:247-271 line by line: ``scale`` is NOT a
    parameter (the reference ignores it on this path, :228,:244,:255), the
    causal mask is top-left aligned ``triu(k=1)`` with fill ``-1e9`` (:261-263)
    and both contractions are ``np.einsum`` with default ``optimize=False``.
    """
    batch, heads, seq_q, head_dim = q.shape
    _, _, seq_k, _ = k.shape
    scale = 1.0 / math.sqrt(head_dim)                               # :253
    scores = np.einsum('bhqd,bhkd->bhqk', q, k) * scale             # :255-258
    if causal:
        mask = np.triu(np.ones((seq_q, seq_k)), k=1).astype(bool)   # :261-262
        scores = np.where(mask, -1e9, scores)                       # :263
    scores_max = scores.max(axis=-1, keepdims=True)                 # :266
    exp_scores = np.exp(scores - scores_max)                        # :267
    attn_weights = exp_scores / exp_scores.sum(axis=-1, keepdims=True)  # :268
    return np.einsum('bhqk,bhkd->bhqd', attn_weights, v)            # :271


# --------------------------------------------------------------------------
# 2. Extended oracle (what the GPU path is compared with).
# --------------------------------------------------------------------------
def _expand_kv(x, hq):
    """GQA/MQA broadcast: kv_head = q_head // (Hq/Hkv)
    (triton_flash.py:95-96, attention_f32.comp:65-67) == repeat_interleave
    along the head axis (tests/test_gqa_unit.py:46-47)."""
    hkv = x.shape[1]
    if hq == hkv:
        return x
    if hq % hkv != 0:
        raise ValueError(f"heads_q ({hq}) must be divisible by heads_kv ({hkv}) for GQA")
    return np.repeat(x, hq // hkv, axis=1)


def _mask(seq_q, seq_k, causal, window, row0=0):
    """Boolean [rows, seq_k] 'masked-out' matrix. Causal is TOP-LEFT aligned:
    key j is visible to query i iff j <= i (triton_flash.py:187,
    attention_f32.comp:169, __init__.py:262, attention_ref.zig:120).
    Sliding window, the Vulkan shader's convention (attention_f32.comp:173-183): causal keeps i - j < window,
    bidirectional keeps |i - j| <= window // 2.  (The reference's generic Triton kernel keeps i - j <= W and
    |i - j| <= W, triton_flash.py:190-194: its W is window = W + 1 causal, window = 2 W bidirectional here.)"""
    i = np.arange(row0, row0 + seq_q)[:, None]
    j = np.arange(seq_k)[None, :]
    m = np.zeros((seq_q, seq_k), dtype=bool)
    if causal:
        m |= j > i
    if window is not None and window > 0:
        if causal:
            m |= (i - j) >= window
        else:
            m |= np.abs(i - j) > (window // 2)
    return m


def attention_ref(q, k, v, causal=True, scale=None, window=-1, acc=np.float64):
    """Returns (O, LSE). O has q's head count; LSE = m + ln(l) over SCALED
    scores, natural log, shape [B,Hq,Sq] (triton_flash.py:232,:437).
    Inputs are used at the precision given (round to bf16 BEFORE calling)."""
    B, Hq, Sq, D = q.shape
    Sk = k.shape[2]
    if scale is None or scale <= 0:
        scale = 1.0 / math.sqrt(D)
    qa = q.astype(acc)
    ka = _expand_kv(k, Hq).astype(acc)
    va = _expand_kv(v, Hq).astype(acc)
    s = np.einsum('bhqd,bhkd->bhqk', qa, ka, optimize=True) * acc(scale)
    m = _mask(Sq, Sk, causal, window)
    if m.any():
        s = np.where(m, -np.inf, s)
    smax = s.max(axis=-1, keepdims=True)
    p = np.exp(s - smax)
    l = p.sum(axis=-1, keepdims=True)
    o = np.einsum('bhqk,bhkd->bhqd', p / l, va, optimize=True)
    lse = (smax + np.log(l))[..., 0]
    return o, lse


def attention_rows(q, k, v, b, h, row0, nrows, causal=True, scale=None, acc=np.float64):
    """O[b,h,row0:row0+nrows,:] and LSE for one (batch, q-head, row block):
    O[rows] = softmax(scale * Q[rows] K^T + mask[rows]) V  (SURVEY 8c)."""
    B, Hq, Sq, D = q.shape
    Hkv, Sk = k.shape[1], k.shape[2]
    if scale is None or scale <= 0:
        scale = 1.0 / math.sqrt(D)
    hk = h // (Hq // Hkv)
    qa = np.asarray(q[b, h, row0:row0 + nrows]).astype(acc)
    ka = np.asarray(k[b, hk]).astype(acc)
    va = np.asarray(v[b, hk]).astype(acc)
    s = (qa @ ka.T) * acc(scale)
    if causal:
        s = np.where(_mask(qa.shape[0], Sk, True, -1, row0=row0), -np.inf, s)
    smax = s.max(axis=-1, keepdims=True)
    p = np.exp(s - smax)
    l = p.sum(axis=-1, keepdims=True)
    return (p / l) @ va, (smax + np.log(l))[:, 0]


def attention_bwd_ref(q, k, v, do, causal=True, scale=None, acc=np.float64, window=-1):
    """Analytic gradients (triton_flash.py:321-347, attention_backward_f32.comp:142-222):
        P = exp(S - LSE); Delta_i = sum_d O_id dO_id; dV = P^T dO; dP = dO V^T;
        dS = P o (dP - Delta) * scale; dQ = dS K; dK = dS^T Q;
    GQA: dK/dV summed over the q-heads of a group.  Returns (dq, dk, dv, o, lse)."""
    B, Hq, Sq, D = q.shape
    Hkv, Sk = k.shape[1], k.shape[2]
    if scale is None or scale <= 0:
        scale = 1.0 / math.sqrt(D)
    g = Hq // Hkv
    qa = q.astype(acc)
    ka = _expand_kv(k, Hq).astype(acc)
    va = _expand_kv(v, Hq).astype(acc)
    doa = do.astype(acc)
    s = np.einsum('bhqd,bhkd->bhqk', qa, ka, optimize=True) * acc(scale)
    m = _mask(Sq, Sk, causal, window)
    if m.any():
        s = np.where(m, -np.inf, s)
    smax = s.max(axis=-1, keepdims=True)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.exp(s - np.where(np.isfinite(smax), smax, 0.0))
        l = e.sum(axis=-1, keepdims=True)
        p = np.where(l > 0, e / np.where(l > 0, l, 1.0), 0.0)     # rows without a visible key contribute nothing
    o = np.einsum('bhqk,bhkd->bhqd', p, va, optimize=True)
    with np.errstate(divide="ignore"):
        lse = (smax + np.log(l))[..., 0]
    delta = (o * doa).sum(axis=-1, keepdims=True)
    dv_full = np.einsum('bhqk,bhqd->bhkd', p, doa, optimize=True)
    dp = np.einsum('bhqd,bhkd->bhqk', doa, va, optimize=True)
    ds = p * (dp - delta) * acc(scale)
    dq = np.einsum('bhqk,bhkd->bhqd', ds, ka, optimize=True)
    dk_full = np.einsum('bhqk,bhqd->bhkd', ds, qa, optimize=True)
    dk = dk_full.reshape(B, Hkv, g, Sk, D).sum(axis=2)
    dv = dv_full.reshape(B, Hkv, g, Sk, D).sum(axis=2)
    return dq, dk, dv, o, lse


def rope_ref(x, cos, sin, interleaved=False):
    """RoPE in fp64. x: [B,H,S,D], cos/sin: [>=S, D/2].
    Half-split (default), restating apply_rope_separate (python/aule/triton_flash.py:680-703):
        x_rot = x * [cos,cos] + rotate_half(x) * [sin,sin], rotate_half(x) = [-x2, x1].
    interleaved=True, restating the Vulkan shader (shaders/attention_f32.comp:98-111) and the torch formula of the
    reference's own test (tests/test_rope_unit.py:76-85): pairs (2i, 2i+1),
        out[2i] = x[2i] cos_i - x[2i+1] sin_i,  out[2i+1] = x[2i] sin_i + x[2i+1] cos_i."""
    x = np.asarray(x, np.float64)
    S, D = x.shape[2], x.shape[3]
    cos = np.asarray(cos, np.float64).reshape(-1, D // 2)[:S]
    sin = np.asarray(sin, np.float64).reshape(-1, D // 2)[:S]
    if interleaved:
        x1, x2 = x[..., 0::2], x[..., 1::2]
        out = np.empty_like(x)
        out[..., 0::2] = x1 * cos[None, None] - x2 * sin[None, None]
        out[..., 1::2] = x1 * sin[None, None] + x2 * cos[None, None]
        return out
    c = np.concatenate([cos, cos], axis=-1)[None, None]
    s = np.concatenate([sin, sin], axis=-1)[None, None]
    x1, x2 = x[..., :D // 2], x[..., D // 2:]
    return x * c + np.concatenate([-x2, x1], axis=-1) * s


def allclose_zig(got, exp, atol, rtol, floor=1e-6):
    """The reference's comparison (tests/test_attention.zig:60-77 compareArrays): an element passes when
    |got-exp| < atol OR |got-exp| / max(|exp|, floor) < rtol.  Returns (ok, number of failing elements, worst abs, worst rel)."""
    got, exp = np.asarray(got, np.float64), np.asarray(exp, np.float64)
    ad = np.abs(got - exp)
    rd = ad / np.maximum(np.abs(exp), floor)
    bad = ~((ad < atol) | (rd < rtol))
    return (not bad.any()), int(bad.sum()), float(ad.max()), float(rd[ad >= atol].max() if (ad >= atol).any() else 0.0)


# --------------------------------------------------------------------------
# 3. Comparison helpers in the reference's tolerance SHAPE.
#    tests/test_attention.zig:60-77 : max_abs < atol OR max_rel < rtol,
#    relative error with a denominator floor.
# --------------------------------------------------------------------------
def paged_decode_ref(q, k_cache, v_cache, block_tables, context_lens, scale=None, window=-1, acc=np.float64):
    """Paged-KV decode, restating the reference's decode kernel python/aule/triton_flash_amd.py:544-660
    (argument meaning of the wrapper :662-740) as a gather + dense softmax:
      q [B,Hq,D]; k_cache/v_cache [num_blocks, block_size, Hkv, D]; block_tables [B,max_blocks] int; context_lens [B] int
      token j of sequence b lives in page block_tables[b, j // block_size] at slot j % block_size   (:597-605)
      kv head of q head h is h // (Hq // Hkv)                                                      (:573-575)
      valid tokens: j < context_len, and with window > 0 also (context_len - 1 - j) < window        (:613-621)
    The reference's online-softmax loop over blocks is algebraically this softmax.  A sequence with
    context_len == 0 returns zeros here (the reference computes 0/0)."""
    q = np.asarray(q)
    B, Hq, D = q.shape
    _, bs, Hkv, _ = k_cache.shape
    g = Hq // Hkv
    if scale is None:
        scale = 1.0 / np.sqrt(D)
    out = np.zeros((B, Hq, D), dtype=acc)
    for b in range(B):
        n = int(context_lens[b])
        if n <= 0:
            continue
        pos = np.arange(n)
        pages = np.asarray(block_tables[b])[pos // bs]
        k = np.asarray(k_cache)[pages, pos % bs].astype(acc)          # [n, Hkv, D]
        v = np.asarray(v_cache)[pages, pos % bs].astype(acc)
        keep = np.ones(n, dtype=bool) if window <= 0 else (n - 1 - pos) < window
        for h in range(Hq):
            s = (k[:, h // g, :] @ q[b, h].astype(acc)) * scale
            s = np.where(keep, s, -np.inf)
            p = np.exp(s - s.max())
            out[b, h] = (p / p.sum()) @ v[:, h // g, :]
    return out


def max_abs_diff(a, b):
    """src/attention_ref.zig:189-197"""
    return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def mean_abs_diff(a, b):
    """src/attention_ref.zig:200-206"""
    return float(np.mean(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))))


def max_rel_diff(actual, expected, floor=1e-6):
    """tests/test_attention.zig:60-77: |a-e| / max(|e|, floor)."""
    a = np.asarray(actual, np.float64)
    e = np.asarray(expected, np.float64)
    return float(np.max(np.abs(a - e) / np.maximum(np.abs(e), floor)))


def rel_err_to_scale(actual, expected):
    """max|a-e| / max|e| : the scale-relative error used for the bf16 gate
    (north_star '<=1e-2 rel bf16'; SURVEY 8d parity gate)."""
    a = np.asarray(actual, np.float64)
    e = np.asarray(expected, np.float64)
    return float(np.max(np.abs(a - e)) / max(float(np.max(np.abs(e))), 1e-30))


def bf16_round(x):
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (what the GPU path sees)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return r.view(np.float32).reshape(x.shape)


def to_bf16_bits(x):
    """fp32 -> uint16 bf16 bit patterns (RNE)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    return ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16).reshape(x.shape)


def from_bf16_bits(b):
    return (np.ascontiguousarray(b, dtype=np.uint16).astype(np.uint32) << 16).view(np.float32).reshape(b.shape)
