"""LayerNorm and BatchNorm written out by hand (forward + backward), checked against autograd and,
on a B200, against this package's fused LayerNorm kernel (reference: explore/understand_ops/*.py).

LayerNorm, per row x in R^H:   mu = mean(x), rstd = 1/sqrt(var(x)+eps), xhat = (x-mu)*rstd
    y  = xhat * gamma + beta
    dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),   g = dy * gamma
    dgamma = sum_rows(dy * xhat),  dbeta = sum_rows(dy)
BatchNorm is the same algebra with the statistics taken over the batch axis instead of the
feature axis (plus running averages at inference).

    python examples/understand_ops/norm_from_scratch.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch


def layernorm_fwd(x, gamma, beta, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)
    xhat = (x - mu) * rstd
    return xhat * gamma + beta, (xhat, rstd)


def layernorm_bwd(dy, gamma, cache):
    xhat, rstd = cache
    g = dy * gamma
    dx = rstd * (g - g.mean(-1, keepdim=True) - xhat * (g * xhat).mean(-1, keepdim=True))
    return dx, (dy * xhat).sum(0), dy.sum(0)


def batchnorm_fwd(x, gamma, beta, eps=1e-5):
    mu = x.mean(0, keepdim=True)
    rstd = torch.rsqrt(x.var(0, unbiased=False, keepdim=True) + eps)
    xhat = (x - mu) * rstd
    return xhat * gamma + beta, (xhat, rstd)


def batchnorm_bwd(dy, gamma, cache):
    xhat, rstd = cache
    g = dy * gamma
    dx = rstd * (g - g.mean(0, keepdim=True) - xhat * (g * xhat).mean(0, keepdim=True))
    return dx, (dy * xhat).sum(0), dy.sum(0)


def main():
    torch.manual_seed(0)
    rows, H = 64, 768
    x = torch.randn(rows, H, dtype=torch.float64, requires_grad=True)
    gamma = torch.randn(H, dtype=torch.float64, requires_grad=True)
    beta = torch.randn(H, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(rows, H, dtype=torch.float64)

    for name, fwd, bwd, ref in (
            ("LayerNorm", layernorm_fwd, layernorm_bwd,
             lambda: torch.nn.functional.layer_norm(x, (H,), gamma, beta, 1e-5)),
            ("BatchNorm", batchnorm_fwd, batchnorm_bwd,
             lambda: torch.nn.functional.batch_norm(x, None, None, gamma, beta, True, 0.0, 1e-5))):
        y, cache = fwd(x.detach(), gamma.detach(), beta.detach())
        dx, dg, db = bwd(dy, gamma.detach(), cache)
        yr = ref()
        gx, gg, gb = torch.autograd.grad(yr, (x, gamma, beta), dy)
        errs = [float((a.detach() - b.detach()).abs().max()) for a, b in ((y, yr), (dx, gx), (dg, gg), (db, gb))]
        print(f"{name}: max abs err  y {errs[0]:.1e}  dx {errs[1]:.1e}  dgamma {errs[2]:.1e}  dbeta {errs[3]:.1e}")
        assert max(errs) < 1e-9

    if torch.cuda.is_available():
        from torchdistpackage_b200.ops import fused
        xb = x.detach().float().cuda().to(torch.bfloat16).requires_grad_(True)
        gb_, bb_ = gamma.detach().float().cuda().to(torch.bfloat16), beta.detach().float().cuda().to(torch.bfloat16)
        yk = fused.layer_norm(xb, gb_, bb_, 1e-5)
        yh, _ = layernorm_fwd(xb.detach().float(), gb_.float(), bb_.float())
        print(f"fused sm_100a LayerNorm kernel vs hand-written: max abs err "
              f"{float((yk.float() - yh).abs().max()):.2e} (bf16 output)")


if __name__ == "__main__":
    main()
