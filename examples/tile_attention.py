"""Tiled (flash-style) attention forward + backward in plain PyTorch with online softmax, checked
against scaled_dot_product_attention (reference study: explore/flash-attn/tile_attn.py).

Forward keeps a running row max `m` and row sum `l` per query tile and rescales the partial
output when the max moves; backward recomputes P tile by tile from the saved log-sum-exp and uses
D = rowsum(dO * O).  This is the state a ring / context-parallel attention step would exchange,
and the algorithm a tcgen05 attention kernel keeps in TMEM."""
import math
import torch
import torch.nn.functional as F


def tiled_attention_fwd(q, k, v, bq=64, bk=64, causal=False):
    B, H, N, D = q.shape
    scale = 1.0 / math.sqrt(D)
    o = torch.zeros_like(q, dtype=torch.float32)
    lse = torch.empty(B, H, N, dtype=torch.float32, device=q.device)
    for i in range(0, N, bq):
        qi = q[:, :, i:i + bq].float() * scale
        m = torch.full((B, H, qi.shape[2]), -float("inf"), device=q.device)
        l = torch.zeros_like(m)
        acc = torch.zeros(B, H, qi.shape[2], D, device=q.device)
        for j in range(0, N, bk):
            if causal and j > i + bq - 1:
                break
            s = qi @ k[:, :, j:j + bk].float().transpose(-1, -2)
            if causal:
                qi_idx = torch.arange(i, i + qi.shape[2], device=q.device)[:, None]
                kj_idx = torch.arange(j, j + s.shape[-1], device=q.device)[None, :]
                s = s.masked_fill(kj_idx > qi_idx, -float("inf"))
            m_new = torch.maximum(m, s.amax(-1))
            p = torch.exp(s - m_new[..., None])
            alpha = torch.exp(m - m_new)
            l = l * alpha + p.sum(-1)
            acc = acc * alpha[..., None] + p @ v[:, :, j:j + bk].float()
            m = m_new
        o[:, :, i:i + bq] = acc / l[..., None]
        lse[:, :, i:i + bq] = m + torch.log(l)
    return o.to(q.dtype), lse


def tiled_attention_bwd(q, k, v, o, lse, do, bq=64, bk=64, causal=False):
    B, H, N, D = q.shape
    scale = 1.0 / math.sqrt(D)
    dq, dk, dv = (torch.zeros_like(t, dtype=torch.float32) for t in (q, k, v))
    delta = (do.float() * o.float()).sum(-1)                    # D_i = rowsum(dO * O)
    for j in range(0, N, bk):
        kj, vj = k[:, :, j:j + bk].float(), v[:, :, j:j + bk].float()
        for i in range(0, N, bq):
            if causal and j > i + bq - 1:
                continue
            qi, doi = q[:, :, i:i + bq].float(), do[:, :, i:i + bq].float()
            s = (qi * scale) @ kj.transpose(-1, -2)
            if causal:
                qi_idx = torch.arange(i, i + qi.shape[2], device=q.device)[:, None]
                kj_idx = torch.arange(j, j + s.shape[-1], device=q.device)[None, :]
                s = s.masked_fill(kj_idx > qi_idx, -float("inf"))
            p = torch.exp(s - lse[:, :, i:i + bq, None])
            dv[:, :, j:j + bk] += p.transpose(-1, -2) @ doi
            dp = doi @ vj.transpose(-1, -2)
            ds = p * (dp - delta[:, :, i:i + bq, None])
            dq[:, :, i:i + bq] += ds @ kj * scale
            dk[:, :, j:j + bk] += ds.transpose(-1, -2) @ qi * scale
    return dq.to(q.dtype), dk.to(k.dtype), dv.to(v.dtype)


if __name__ == "__main__":
    torch.manual_seed(0)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    for causal in (False, True):
        q, k, v = (torch.randn(2, 4, 256, 64, device=dev, requires_grad=True) for _ in range(3))
        ref = F.scaled_dot_product_attention(q, k, v, is_causal=causal)
        do = torch.randn_like(ref)
        gq, gk, gv = torch.autograd.grad(ref, (q, k, v), do)
        o, lse = tiled_attention_fwd(q.detach(), k.detach(), v.detach(), causal=causal)
        dq, dk, dv = tiled_attention_bwd(q.detach(), k.detach(), v.detach(), o, lse, do, causal=causal)
        for a, b, n in ((o, ref, "o"), (dq, gq, "dq"), (dk, gk, "dk"), (dv, gv, "dv")):
            err = (a - b).abs().max().item()
            assert err < 2e-4, (causal, n, err)
        print(f"causal={causal}: tiled attention fwd/bwd match SDPA")
