"""CPU emulation of the data flow of the tcgen05 attention kernels (csrc/attn/*.cu): the same tile
loops, masks, (lazy) online-softmax update order, scale placement and bf16 rounding points, written
with torch ops and checked against a dense fp32 reference.  It pins down the *algorithm* the kernels
implement (what is accumulated where, in which order); the hardware-specific parts are checked by
scripts/attn_check.py on a B200."""
import math

import pytest
import torch

T_Q = T_KV = 128
LOG2E = 1.4426950408889634


def _bf16(x):
    return x.to(torch.bfloat16).float()


RESCALE_THRESHOLD = 8.0        # log2 units, kRescaleThreshold of attn_fwd_sm100.cu


def emulate_forward(q, k, v, causal, scale, stats=None):
    """q, k, v: [T, 64] fp32 holding bf16 values (one batch / head).  Mirrors attn_fwd_sm100.cu:
    O accumulates in TMEM across key tiles; the running max is only advanced -- and O / l rescaled
    -- when the tile's max exceeds it by more than 2^8 (decided per warp = 32 rows), so P is
    computed against a possibly stale max (values up to 2^8, still exact in bf16 / fp32)."""
    T = q.shape[0]
    out = torch.zeros(T, 64)
    lse = torch.zeros(T)
    scale_log2 = scale * LOG2E
    for q0 in range(0, T, 2 * T_Q):
        for wg in range(2):
            r0 = q0 + wg * T_Q
            if r0 >= T:
                continue
            n_kv = (q0 // T_KV + 1 + wg) if causal else T // T_KV
            rows = torch.arange(T_Q)
            o = torch.zeros(T_Q, 64)                                          # TMEM accumulator
            m_run = torch.zeros(T_Q)
            l_run = torch.zeros(T_Q)
            for j in range(n_kv):
                s = q[r0:r0 + T_Q] @ k[j * T_KV:(j + 1) * T_KV].t()            # UMMA, fp32 acc
                if causal and j == n_kv - 1:
                    s = torch.where(torch.arange(T_KV)[None, :] <= rows[:, None], s,
                                    torch.full_like(s, -math.inf))
                sm = s.max(-1).values * scale_log2
                if j == 0:
                    m_run = sm.clone()
                else:
                    need = (sm > m_run + RESCALE_THRESHOLD).view(4, 32).any(-1)   # warp vote
                    if stats is not None:
                        stats["rescales"] = stats.get("rescales", 0) + int(need.sum())
                    need_rows = need.repeat_interleave(32)
                    m_new = torch.where(need_rows, torch.maximum(m_run, sm), m_run)
                    alpha = torch.exp2(m_run - m_new)
                    o = o * alpha[:, None]                                    # tcgen05.ld / st
                    l_run = l_run * alpha
                    m_run = m_new
                p = torch.exp2(s * scale_log2 - m_run[:, None])               # exp2(-inf) = 0
                l_run = l_run + p.sum(-1)
                o = o + _bf16(p) @ v[j * T_KV:(j + 1) * T_KV]                 # P in bf16, fp32 acc
            out[r0:r0 + T_Q] = _bf16(o / l_run[:, None])
            lse[r0:r0 + T_Q] = (m_run + torch.log2(l_run)) * math.log(2.0)
    return out, lse


SUB = 64                       # keys per step of the dQ kernel


def emulate_backward(q, k, v, o, do, lse, causal, scale):
    """Mirrors the two backward kernels: attn_bwd_dq_sm100.cu (query tile stationary, 64-key
    sub-tiles, delta computed from the bf16 dO and O tiles, scale applied to dQ at the end) and
    attn_bwd_sm100.cu (key tile stationary, scale applied to dK at the end); dS is stored as bf16
    WITHOUT the softmax scale in both."""
    T = q.shape[0]
    n = T // T_KV
    scale_log2 = scale * LOG2E
    delta = (do * o).sum(-1)                                                  # fp32 from bf16 tiles
    # ---- kernel 1: dQ
    dq = torch.zeros(T, 64)
    for i in range(n):
        qi, doi = q[i * T_Q:(i + 1) * T_Q], do[i * T_Q:(i + 1) * T_Q]
        lse2 = lse[i * T_Q:(i + 1) * T_Q] * LOG2E
        acc = torch.zeros(T_Q, 64)
        n_sub = (2 * i + 2) if causal else T // SUB
        for j in range(n_sub):
            kj, vj = k[j * SUB:(j + 1) * SUB], v[j * SUB:(j + 1) * SUB]
            s = qi @ kj.t()
            dp = doi @ vj.t()
            p = torch.exp2(s * scale_log2 - lse2[:, None])
            if causal:
                t = i * T_Q + torch.arange(T_Q)[:, None]
                p = torch.where(j * SUB + torch.arange(SUB)[None, :] <= t, p, torch.zeros_like(p))
            ds16 = _bf16(p * (dp - delta[i * T_Q:(i + 1) * T_Q, None]))
            acc += ds16 @ kj
        dq[i * T_Q:(i + 1) * T_Q] = _bf16(acc * scale)
    # ---- kernel 2: dK, dV
    dk = torch.zeros(T, 64)
    dv = torch.zeros(T, 64)
    for j in range(n):
        kj, vj = k[j * T_KV:(j + 1) * T_KV], v[j * T_KV:(j + 1) * T_KV]
        dk_acc = torch.zeros(T_KV, 64)
        dv_acc = torch.zeros(T_KV, 64)
        for i in range(j if causal else 0, n):
            qi, doi = q[i * T_Q:(i + 1) * T_Q], do[i * T_Q:(i + 1) * T_Q]
            s = qi @ kj.t()
            dp = doi @ vj.t()
            lse2 = lse[i * T_Q:(i + 1) * T_Q] * LOG2E
            p = torch.exp2(s * scale_log2 - lse2[:, None])
            if causal and i == j:
                keep = torch.arange(T_KV)[None, :] <= torch.arange(T_Q)[:, None]
                p = torch.where(keep, p, torch.zeros_like(p))
            ds = p * (dp - delta[i * T_Q:(i + 1) * T_Q, None])
            dv_acc += _bf16(p).t() @ doi
            dk_acc += _bf16(ds).t() @ qi
        dk[j * T_KV:(j + 1) * T_KV] = _bf16(dk_acc * scale)
        dv[j * T_KV:(j + 1) * T_KV] = _bf16(dv_acc)
    return dq, dk, dv


def test_lazy_rescale_with_growing_scores():
    """Scores that keep growing along the key axis force the thresholded rescale path (and stale
    maxima of up to 2^8 in between); the result must still match the dense softmax."""
    torch.manual_seed(5)
    T = 512
    scale = 64 ** -0.5
    q = _bf16(torch.randn(T, 64))
    k = _bf16(torch.randn(T, 64) * 0.3 + torch.linspace(0, 6, T)[:, None] * q.mean(0)[None, :].sign())
    v = _bf16(torch.randn(T, 64))
    stats = {}
    out, lse = emulate_forward(q, k, v, False, scale * 8, stats)
    s = (q @ k.t()) * scale * 8
    ref = torch.softmax(s, -1) @ v
    assert stats.get("rescales", 0) > 0
    assert (out - ref).abs().max() / ref.abs().max() < 1.5e-2
    assert (lse - torch.logsumexp(s, -1)).abs().max() < 2e-3


@pytest.mark.parametrize("T,causal", [(128, True), (256, True), (384, True), (256, False), (512, False)])
def test_attention_kernel_dataflow_matches_dense_reference(T, causal):
    torch.manual_seed(T + int(causal))
    scale = 64 ** -0.5
    q, k, v = (_bf16(torch.randn(T, 64) * 0.8) for _ in range(3))
    do = _bf16(torch.randn(T, 64) * 0.5)

    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    s = (qr @ kr.t()) * scale
    if causal:
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool).triu(1), -math.inf)
    ref_lse = torch.logsumexp(s, -1)
    ref_o = torch.softmax(s, -1) @ vr
    ref_o.backward(do)

    out, lse = emulate_forward(q, k, v, causal, scale)
    assert (out - ref_o.detach()).abs().max() / ref_o.abs().max() < 1.5e-2
    assert (lse - ref_lse.detach()).abs().max() < 1e-3

    dq, dk, dv = emulate_backward(q, k, v, out, do, lse, causal, scale)
    for got, ref in ((dq, qr.grad), (dk, kr.grad), (dv, vr.grad)):
        assert (got - ref).abs().max() / ref.abs().max() < 2e-2
