"""1-GPU numerics of the fused elementwise / optimizer / norm / loss kernels vs fp32 PyTorch."""
import json, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torchdistpackage_b200._C as C

torch.manual_seed(0)
dev = "cuda"
res = {}

def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()

# ---- AdamW (bf16 param + fp32 master, bf16 grad) vs torch.optim.AdamW on fp32
n = 1_000_003
p32 = torch.randn(n, device=dev)
ref_p = p32.clone().requires_grad_(True)
opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.1)
master = p32.clone()
pb = p32.to(torch.bfloat16)
m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
copy_out = torch.empty(n, device=dev, dtype=torch.bfloat16)
for step in range(1, 4):
    g = torch.randn(n, device=dev).to(torch.bfloat16)
    ref_p.grad = g.float()
    opt.step()
    C.adamw(pb, master, g, m, v, 1e-3, 0.9, 0.95, 1e-8, 0.1, step, True, 1.0, None, copy_out)
res["adamw_master_rel"] = rel(master, ref_p.detach())
res["adamw_bf16_rel"] = rel(pb, ref_p.detach())
res["adamw_copy_equal"] = bool(torch.equal(copy_out, pb))
# Adam (L2) fp32 params, fp32 grads, no master
p2 = torch.randn(4099, device=dev); r2 = p2.clone().requires_grad_(True)
o2 = torch.optim.Adam([r2], lr=1e-2, weight_decay=0.01)
m2 = torch.zeros_like(p2); v2 = torch.zeros_like(p2)
for step in range(1, 4):
    g = torch.randn_like(p2); r2.grad = g.clone(); o2.step()
    C.adamw(p2, None, g, m2, v2, 1e-2, 0.9, 0.999, 1e-8, 0.01, step, False)
res["adam_fp32_rel"] = rel(p2, r2.detach())

# ---- EMA
ema = torch.randn(n, device=dev); par = torch.randn(n, device=dev)
ref = ema * 0.999 + par * 0.001
C.ema_update(ema, par, 0.999)
res["ema_rel"] = rel(ema, ref)
ts = [torch.randn(s, device=dev) for s in (5, 1024, 77777)]
es = [torch.randn_like(t) for t in ts]
refs = [e * 0.9 + t * 0.1 for e, t in zip(es, ts)]
tab = lambda xs: torch.tensor([x.data_ptr() for x in xs], dtype=torch.int64, device=dev)
C.ema_update_multi(tab(es), tab(ts), torch.tensor([t.numel() for t in ts], dtype=torch.int64, device=dev),
                   torch.ones(3, dtype=torch.int32, device=dev), torch.ones(3, dtype=torch.int32, device=dev), 0.9)
res["ema_multi_rel"] = max(rel(e, r) for e, r in zip(es, refs))

# ---- sumsq / scale / cast
x = torch.randn(3_000_001, device=dev, dtype=torch.bfloat16)
out = torch.zeros(1, device=dev)
C.sumsq(x, out)
res["sumsq_rel"] = abs(out.item() - x.float().pow(2).sum().item()) / x.float().pow(2).sum().item()
y = x.clone(); C.scale_(y, 0.5, None)
res["scale_rel"] = rel(y, x.float() * 0.5)
d = torch.empty(x.numel(), device=dev); C.cast_copy(d, x, 2.0)
res["cast_rel"] = rel(d, x.float() * 2)

# ---- LayerNorm fwd/bwd (+ residual)
for cols in (768, 1024, 4096):
    rows = 4100
    xx = torch.randn(rows, cols, device=dev, dtype=torch.bfloat16)
    rr = torch.randn(rows, cols, device=dev, dtype=torch.bfloat16)
    gam = (torch.randn(cols, device=dev) * 0.1 + 1).to(torch.bfloat16)
    bet = (torch.randn(cols, device=dev) * 0.1).to(torch.bfloat16)
    yy = torch.empty_like(xx); ro = torch.empty_like(xx)
    mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
    C.layernorm_fwd(xx, rr, gam, bet, yy, ro, mean, rstd, 1e-5)
    s = (xx.float() + rr.float())
    s_b = s.to(torch.bfloat16).float().requires_grad_(True)
    g32 = gam.float().requires_grad_(True); b32 = bet.float().requires_grad_(True)
    yref = F.layer_norm(s_b, (cols,), g32, b32, 1e-5)
    res[f"ln_fwd_rel_{cols}"] = rel(yy, yref)
    res[f"ln_resid_rel_{cols}"] = rel(ro, s)
    dy = torch.randn(rows, cols, device=dev, dtype=torch.bfloat16)
    dres = torch.randn(rows, cols, device=dev, dtype=torch.bfloat16)
    yref.backward(dy.float())
    dx = torch.empty_like(xx); dg = torch.empty(cols, device=dev); db = torch.empty(cols, device=dev)
    C.layernorm_bwd(dy, ro, gam, mean, rstd, dx, dres, dg, db)
    res[f"ln_dx_rel_{cols}"] = rel(dx, s_b.grad + dres.float())
    res[f"ln_dgamma_rel_{cols}"] = rel(dg, g32.grad)
    res[f"ln_dbeta_rel_{cols}"] = rel(db, b32.grad)

# ---- colsum
xx = torch.randn(5000, 3072, device=dev, dtype=torch.bfloat16)
o = torch.empty(3072, device=dev)
C.colsum(xx, o)
res["colsum_rel"] = rel(o, xx.float().sum(0))

# ---- cross entropy fwd+bwd
rows, vocab = 2048, 50304
lg = (torch.randn(rows, vocab, device=dev) * 2).to(torch.bfloat16)
tgt = torch.randint(0, vocab, (rows,), device=dev)
tgt[5] = -100
l32 = lg.float().requires_grad_(True)
lref = F.cross_entropy(l32, tgt, reduction="none", ignore_index=-100)
(lref.sum() / rows).backward()
loss = torch.empty(rows, device=dev)
lgc = lg.clone()
C.cross_entropy_fwd_bwd(lgc, tgt, loss, 1.0 / rows, -100)
res["ce_loss_rel"] = rel(loss, lref.detach())
res["ce_grad_rel"] = rel(lgc, l32.grad)

torch.cuda.synchronize()
bad = {k: v for k, v in res.items() if (isinstance(v, float) and not (v < 2e-2)) or v is False}
res["all_ok"] = len(bad) == 0
print(json.dumps(res, indent=1))
print("BAD", bad)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fused_check.json", "w"), indent=1)
