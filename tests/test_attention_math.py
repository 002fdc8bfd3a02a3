"""CPU emulation of the data flow of the tcgen05 attention kernels (csrc/attn/*.cu): the same tile
loops, masks, online-softmax update order, scale placement and bf16 rounding points, written with
torch ops and checked against a dense fp32 reference.  It pins down the *algorithm* the kernels
implement (what is accumulated where, in which order); the hardware-specific parts are checked by
scripts/attn_check.py on a B200."""
import math

import pytest
import torch

T_Q = T_KV = 128
LOG2E = 1.4426950408889634


def _bf16(x):
    return x.to(torch.bfloat16).float()


def emulate_forward(q, k, v, causal, scale):
    """q, k, v: [T, 64] fp32 holding bf16 values (one batch / head).  Mirrors attn_fwd_sm100.cu."""
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
            o = torch.zeros(T_Q, 64)
            m_run = torch.full((T_Q,), -math.inf)
            l_run = torch.zeros(T_Q)
            o_tile = None
            for j in range(n_kv):
                s = q[r0:r0 + T_Q] @ k[j * T_KV:(j + 1) * T_KV].t()            # UMMA, fp32 acc
                diag = causal and j == n_kv - 1
                valid = (torch.arange(T_KV)[None, :] <= rows[:, None]) if diag else \
                    torch.ones(T_Q, T_KV, dtype=torch.bool)
                mx = torch.where(valid, s, torch.full_like(s, -math.inf)).max(-1).values
                m_new = torch.maximum(m_run, mx * scale_log2)
                if j > 0:
                    o = o + o_tile                                            # fold previous P.V
                alpha = torch.exp2(m_run - m_new)
                o = o * alpha[:, None]
                l_run = l_run * alpha
                p = torch.exp2(s * scale_log2 - m_new[:, None])
                p = torch.where(valid, p, torch.zeros_like(p))
                l_run = l_run + p.sum(-1)
                m_run = m_new
                o_tile = _bf16(p) @ v[j * T_KV:(j + 1) * T_KV]                # P in bf16, fp32 acc
            o = o + o_tile
            out[r0:r0 + T_Q] = _bf16(o / l_run[:, None])
            lse[r0:r0 + T_Q] = (m_run + torch.log2(l_run)) * math.log(2.0)
    return out, lse


def emulate_backward(q, k, v, o, do, lse, causal, scale):
    """Mirrors attn_bwd_sm100.cu: one pass per key tile, dQ accumulated in fp32 across key tiles."""
    T = q.shape[0]
    n = T // T_KV
    delta = (do * o).sum(-1)
    dq_acc = torch.zeros(T, 64)
    dk = torch.zeros(T, 64)
    dv = torch.zeros(T, 64)
    scale_log2 = scale * LOG2E
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
            ds = scale * p * (dp - delta[i * T_Q:(i + 1) * T_Q, None])
            p16, ds16 = _bf16(p), _bf16(ds)
            dv_acc += p16.t() @ doi
            dk_acc += ds16.t() @ qi
            dq_acc[i * T_Q:(i + 1) * T_Q] += ds16 @ kj
        dk[j * T_KV:(j + 1) * T_KV] = _bf16(dk_acc)
        dv[j * T_KV:(j + 1) * T_KV] = _bf16(dv_acc)
    return _bf16(dq_acc), dk, dv


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
