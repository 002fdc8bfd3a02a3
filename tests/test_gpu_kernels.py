"""GPU numerics: every native sm_100a kernel against a plain PyTorch fp32 reference of the same
op (run on the B200 box: ``pytest -m gpu``).  These fail loudly if the extension is not loaded."""
import os
import subprocess
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _C():
    import torchdistpackage_b200 as tdp
    return tdp.ops.native(required=True)


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def assert_close(a, ref, rtol=1.6e-2, atol_frac=4e-3, cos=0.9995, what=""):
    """Per-element check: |a - ref| <= rtol * |ref| + atol_frac * max|ref| for every element (a
    bf16 result carries 2^-8 relative rounding plus accumulation-order noise near zero), and the
    cosine similarity of the whole tensors."""
    a, ref = a.float(), ref.float()
    tol = rtol * ref.abs() + atol_frac * ref.abs().max().clamp_min(1e-12)
    bad = (a - ref).abs() > tol
    assert not bad.any(), (what, int(bad.sum()), float(((a - ref).abs() / tol).max()))
    c = torch.nn.functional.cosine_similarity(a.flatten(), ref.flatten(), dim=0).item()
    assert c > cos, (what, c)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("block_n", [128, 256])
def test_gemm_operand_majors(ta, tb, block_n):
    C = _C()
    torch.manual_seed(0)
    M, N, K = 384, 520, 264
    a = torch.randn((K, M) if ta else (M, K), device="cuda", dtype=torch.bfloat16)
    b = torch.randn((N, K) if tb else (K, N), device="cuda", dtype=torch.bfloat16)
    c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    C.gemm(a, b, c, ta, tb, block_n=block_n)
    ref = (a.float().t() if ta else a.float()) @ (b.float().t() if tb else b.float())
    assert rel(c, ref) < 1e-2


def test_gemm_fused_epilogues():
    C = _C()
    torch.manual_seed(1)
    M, N, K = 512, 1024, 512
    a = torch.randn(M, K, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(K, N, device="cuda", dtype=torch.bfloat16) * 0.05
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    res = torch.randn(M, N, device="cuda", dtype=torch.bfloat16)
    z = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    C.gemm(a, b, c, bias=bias, act=1, aux_out=z)
    pre = a.float() @ b.float() + bias.float()
    assert rel(z, pre) < 1e-2 and rel(c, F.gelu(pre, approximate="tanh")) < 1e-2
    C.gemm(a, b, c, bias=bias, residual=res)
    assert rel(c, pre + res.float()) < 1e-2
    zz = z.float().requires_grad_(True)
    dg, = torch.autograd.grad(F.gelu(zz, approximate="tanh").sum(), zz)
    C.gemm(a, b, c, act=3, aux_in=z)
    assert rel(c, (a.float() @ b.float()) * dg) < 1e-2
    c32 = torch.randn(M, N, device="cuda")
    ref = c32 + a.float() @ b.float()
    C.gemm(a, b, c32, accumulate=True)
    assert rel(c32, ref) < 1e-3


def test_gemm_split_k_weight_gradient_shape():
    from torchdistpackage_b200.ops import linear as L
    C = _C()
    torch.manual_seed(7)
    x = torch.randn(8192, 384, device="cuda", dtype=torch.bfloat16)      # [tokens, in]
    dy = torch.randn(8192, 520, device="cuda", dtype=torch.bfloat16)     # [tokens, out]
    ref = x.float().t() @ dy.float()
    acc = torch.zeros(384, 520, device="cuda")
    C.gemm(x, dy, acc, True, False, split_k=5, block_n=128)
    assert rel(acc, ref) < 2e-3
    assert rel(L.gemm(x, dy, trans_a=True), ref) < 1e-2                   # auto split + cast
    assert rel(L.gemm(x, dy, trans_a=True, split_k=1), ref) < 1e-2


def test_gemm_from_autograd_thread_and_linear_module():
    from torchdistpackage_b200.ops import linear as L
    torch.manual_seed(2)
    x = torch.randn(4, 96, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    for layout in ("kn", "nk"):
        w = (torch.randn(256, 512, device="cuda") * 0.05).to(torch.bfloat16)
        wl = (w if layout == "kn" else w.t().contiguous()).requires_grad_(True)
        b = torch.randn(512, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        y = L.linear(x, wl, b, layout=layout)
        g = torch.randn_like(y)
        gx, gw, gb = torch.autograd.grad(y, (x, wl, b), g)        # GEMMs run on the autograd thread
        xf, wf, bf = x.detach().float().requires_grad_(True), w.float().requires_grad_(True), \
            b.detach().float().requires_grad_(True)
        yr = xf @ wf + bf
        rx, rw, rb = torch.autograd.grad(yr, (xf, wf, bf), g.float())
        assert rel(y, yr) < 1e-2 and rel(gx, rx) < 1e-2 and rel(gb, rb) < 1e-2
        assert rel(gw, rw if layout == "kn" else rw.t()) < 1e-2


def test_fused_mlp_matches_torch():
    from torchdistpackage_b200.ops import linear as L
    torch.manual_seed(3)
    x = torch.randn(256, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w1 = (torch.randn(256, 1024, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    b1 = torch.randn(1024, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w2 = (torch.randn(1024, 256, device="cuda") * 0.05).to(torch.bfloat16).requires_grad_(True)
    b2 = torch.randn(256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = L.mlp(x, w1, b1, w2, b2, act="gelu_tanh", residual=x)
    g = torch.randn_like(y)
    grads = torch.autograd.grad(y, (x, w1, b1, w2, b2), g)
    f = [t.detach().float().requires_grad_(True) for t in (x, w1, b1, w2, b2)]
    yr = F.gelu(f[0] @ f[1] + f[2], approximate="tanh") @ f[3] + f[4] + f[0]
    refs = torch.autograd.grad(yr, f, g.float())
    assert rel(y, yr) < 1e-2
    for a, b in zip(grads, refs):
        assert rel(a, b) < 2e-2


def test_layernorm_and_cross_entropy():
    from torchdistpackage_b200.ops import fused
    torch.manual_seed(4)
    x = torch.randn(1000, 768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    r = torch.randn(1000, 768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    w = (torch.rand(768, device="cuda") + 0.5).to(torch.bfloat16).requires_grad_(True)
    b = torch.randn(768, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y, s = fused.layer_norm(x, w, b, 1e-5, residual=r)
    gy, gs = torch.randn_like(y), torch.randn_like(s)
    grads = torch.autograd.grad((y, s), (x, r, w, b), (gy, gs))
    f = [t.detach().float().requires_grad_(True) for t in (x, r, w, b)]
    sr = f[0] + f[1]
    yr = F.layer_norm(sr, (768,), f[2], f[3], 1e-5)
    refs = torch.autograd.grad((yr, sr), f, (gy.float(), gs.float()))
    assert rel(y, yr) < 2e-2 and rel(s, sr) < 1e-2
    for a, bb in zip(grads, refs):
        assert rel(a, bb) < 3e-2
    logits = (torch.randn(512, 50304, device="cuda") * 2).to(torch.bfloat16).requires_grad_(True)
    tgt = torch.randint(0, 50304, (512,), device="cuda")
    lf = logits.detach().float().requires_grad_(True)
    ref = F.cross_entropy(lf, tgt)
    ref.backward()
    loss = fused.cross_entropy(logits.clone().requires_grad_(True) * 1.0, tgt)
    assert abs(loss.item() - ref.item()) / ref.item() < 1e-3
    lg = logits.detach().clone().requires_grad_(True)
    out = fused._CrossEntropyFn.apply(lg, tgt, -100)
    out.backward()
    assert rel(lg.grad, lf.grad) < 2e-2


def test_adamw_ema_norm_kernels():
    C = _C()
    torch.manual_seed(5)
    n = 300_001
    p32 = torch.randn(n, device="cuda")
    refp = p32.clone().requires_grad_(True)
    opt = torch.optim.AdamW([refp], lr=1e-3, betas=(0.9, 0.95), weight_decay=0.1)
    master, pb = p32.clone(), p32.to(torch.bfloat16)
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        g = torch.randn(n, device="cuda").to(torch.bfloat16)
        refp.grad = g.float(); opt.step()
        C.adamw(pb, master, g, m, v, 1e-3, 0.9, 0.95, 1e-8, 0.1, step, True, 1.0, None, None)
    assert rel(master, refp.detach()) < 1e-4 and rel(pb, refp.detach()) < 1e-2
    ema, par = torch.randn(n, device="cuda"), torch.randn(n, device="cuda")
    ref = ema * 0.99 + par * 0.01
    C.ema_update(ema, par, 0.99)
    assert rel(ema, ref) < 1e-5
    x = torch.randn(1_000_003, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(1, device="cuda")
    C.sumsq(x, out)
    assert abs(out.item() / x.float().pow(2).sum().item() - 1) < 1e-3


def test_gpt2_native_step_matches_torch_reference():
    """One optimizer step of the native GPT-2 (tiny) vs the same math written in plain torch."""
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.models.gpt2 import build_gpt2
    torch.manual_seed(6)
    model = build_gpt2("tiny", device="cuda")
    tok = torch.randint(0, model.cfg.vocab_size, (4, model.cfg.seq_len + 1), device="cuda")
    loss = model(tok[:, :-1], tok[:, 1:])
    loss.backward()
    # plain-torch replica in fp32
    P = {n: p.detach().float().requires_grad_(True) for n, p in model.named_parameters()}
    x = P["wte.weight"][tok[:, :-1]] + P["wpe.weight"][: model.cfg.seq_len]
    for i in range(model.cfg.n_layer):
        g = lambda k: P[f"blocks.{i}.{k}"]
        h = F.layer_norm(x, (x.shape[-1],), g("ln_1.weight"), g("ln_1.bias"))
        qkv = h @ g("w_qkv") + g("b_qkv")
        B, T, D = x.shape
        q, k, v = qkv.view(B, T, 3, model.cfg.n_head, -1).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(B, T, D)
        x = x + o @ g("w_proj") + g("b_proj")
        h = F.layer_norm(x, (D,), g("ln_2.weight"), g("ln_2.bias"))
        x = x + F.gelu(h @ g("w_fc1") + g("b_fc1"), approximate="tanh") @ g("w_fc2") + g("b_fc2")
    x = F.layer_norm(x, (x.shape[-1],), P["ln_f.weight"], P["ln_f.bias"])
    ref = F.cross_entropy((x @ P["wte.weight"].t()).view(-1, model.cfg.vocab_size), tok[:, 1:].reshape(-1))
    assert abs(loss.item() - ref.item()) / ref.item() < 5e-3
    # every parameter gradient of the whole model against autograd through the fp32 replica
    ref.backward()
    for n, p in model.named_parameters():
        g, gr = p.grad.float(), P[n].grad
        assert torch.isfinite(g).all(), n
        err = ((g - gr).abs().max() / gr.abs().max().clamp_min(1e-12)).item()
        cosv = F.cosine_similarity(g.flatten(), gr.flatten(), dim=0).item()
        assert err < 4e-2 and cosv > 0.998, (n, err, cosv)


def test_graft_smoke():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.smoke()


@pytest.mark.multigpu
def test_multi_gpu_collectives_and_fused_tp():
    """Runs the 2-GPU numerics scripts (symmetric collectives, fused GEMM+collective, TP block)."""
    n = 2
    for script, port in (("scripts/symm_check.py", 29621), ("scripts/tp_check.py", 29622)):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                            f"--nproc-per-node={n}", "--master-addr", "127.0.0.1", "--master-port",
                            str(port), os.path.join(ROOT, script)], cwd=ROOT, capture_output=True,
                           text=True, timeout=240)
        assert r.returncode == 0 and "ALL_OK True" in r.stdout, (script, r.stdout[-3000:], r.stderr[-3000:])


def test_ddp_direct_weight_gradients_match_autograd():
    """NaiveDDP bucket views + ops.linear.wgrad: weight gradients written straight into the
    bucket (overwrite on the first micro-step, accumulate on the second) equal the gradients
    autograd accumulates for the un-wrapped model."""
    import copy
    import torchdistpackage_b200 as tdp
    from torchdistpackage_b200.models.gpt2 import build_gpt2
    torch.manual_seed(7)
    base = build_gpt2("tiny", device="cuda")
    wrapped = copy.deepcopy(base)
    ddp = tdp.NaiveDDP(wrapped, sync=False, gradient_as_bucket_view=True, num_grad_acc_iter=2)
    direct = [n for n, p in wrapped.named_parameters() if hasattr(p, "_tdp_main_grad")]
    assert any("w_fc1" in n for n in direct)
    toks = [torch.randint(0, base.cfg.vocab_size, (4, base.cfg.seq_len + 1), device="cuda")
            for _ in range(2)]
    for step in range(2):           # the second step checks that `fresh` is re-armed by finalize()
        base.zero_grad(set_to_none=True)
        ddp.zero_grad()
        for tok in toks:
            base(tok[:, :-1], tok[:, 1:]).backward()
            ddp(tok[:, :-1], tok[:, 1:]).backward()
        ddp.reduce_gradients()
        for (n, p), (_, q) in zip(base.named_parameters(), wrapped.named_parameters()):
            ref = p.grad.float()
            err = (q.grad.float() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6)
            assert err < 3e-2, (step, n, err)


@pytest.mark.parametrize("ta,tb", [(False, False), (False, True), (True, False)])
@pytest.mark.parametrize("block_n", [128, 256])
def test_gemm_2cta_pairs(ta, tb, block_n):
    """cta_group::2 kernel (256 x {128|256} tile per CTA pair): operand majors, ragged edges and a
    fused epilogue against the fp32 product."""
    C = _C()
    torch.manual_seed(2)
    M, N, K = 1000, 760, 520
    a = torch.randn((K, M) if ta else (M, K), device="cuda", dtype=torch.bfloat16)
    b = torch.randn((N, K) if tb else (K, N), device="cuda", dtype=torch.bfloat16)
    bias = torch.randn(N, device="cuda", dtype=torch.bfloat16)
    c = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    C.gemm(a, b, c, ta, tb, bias=bias, block_n=block_n, cta_group=2)
    ref = (a.float().t() if ta else a.float()) @ (b.float().t() if tb else b.float()) + bias.float()
    assert rel(c, ref) < 1e-2


def test_gemm_auto_selection_long_k_and_weight_gradient_shapes():
    """ops.linear.gemm picks the CTA-pair kernels for long-K products and 256x128 pair tiles for
    weight-gradient shapes; results must not depend on the kernel that was selected."""
    from torchdistpackage_b200.ops import linear as L
    torch.manual_seed(3)
    x = torch.randn(4096, 768, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(4096, 3072, device="cuda", dtype=torch.bfloat16)
    dw = L.gemm(x, dy, trans_a=True)                       # [768, 3072], K = 4096 tokens
    assert rel(dw, x.float().t() @ dy.float()) < 1e-2
    w = torch.randn(3072, 768, device="cuda", dtype=torch.bfloat16) * 0.05
    y = L.gemm(dy, w)                                      # K = 3072 >= 2048: CTA pairs
    assert rel(y, dy.float() @ w.float()) < 1e-2
    small = L.gemm(x[:, :768], x[:, :768], trans_a=True)   # 768 x 768, K = 4096: stream-K
    assert rel(small, x.float().t() @ x.float()) < 1e-2


# ------------------------------------------------------------------ native tcgen05 attention
def _dense_attention(qkv, H, causal):
    B, T, D3 = qkv.shape
    dh = D3 // 3 // H
    q, k, v = qkv.view(B, T, 3, H, dh).permute(2, 0, 3, 1, 4)
    s = (q @ k.transpose(-1, -2)) * dh ** -0.5
    if causal:
        s = s.masked_fill(torch.ones(T, T, device=qkv.device, dtype=torch.bool).triu(1), float("-inf"))
    lse = torch.logsumexp(s, -1)
    o = torch.softmax(s, -1) @ v
    return o.transpose(1, 2).reshape(B, T, H * dh), lse


@pytest.mark.parametrize("T", [128, 384, 1024])
@pytest.mark.parametrize("causal", [True, False])
def test_native_attention_forward_and_all_gradients(T, causal):
    """csrc/attn: forward (output + LSE) and dq / dk / dv of the three tcgen05 kernels against an
    fp32 dense softmax(QK^T)V reference and autograd through it."""
    os.environ["TDP_ATTN"] = "native"
    from torchdistpackage_b200.ops.attention import native_attention_forward, packed_attention
    torch.manual_seed(11 + T)
    B, H = (3, 5) if T < 1024 else (2, 12)          # enough work items for several persistent rounds
    qkv = (torch.randn(B, T, 3 * H * 64, device="cuda") * 0.8).to(torch.bfloat16)
    out, lse = native_attention_forward(qkv, H, causal, return_lse=True)
    ro, rl = _dense_attention(qkv.float(), H, causal)
    assert_close(out, ro, rtol=2e-2, atol_frac=6e-3, what="attention output")
    assert (lse - rl).abs().max().item() < 5e-3
    qg = qkv.clone().requires_grad_(True)
    dout = (torch.randn(B, T, H * 64, device="cuda") * 0.5).to(torch.bfloat16)
    packed_attention(qg, H, causal).backward(dout)
    qr = qkv.float().requires_grad_(True)
    _dense_attention(qr, H, causal)[0].backward(dout.float())
    g, gr = qg.grad.float().view(B, T, 3, H, 64), qr.grad.view(B, T, 3, H, 64)
    for i, name in enumerate(("dq", "dk", "dv")):
        assert_close(g[:, :, i], gr[:, :, i], rtol=3e-2, atol_frac=1e-2, cos=0.999, what=name)


def test_native_attention_is_deterministic():
    """No atomics anywhere in the attention backward: two runs are bitwise identical."""
    os.environ["TDP_ATTN"] = "native"
    from torchdistpackage_b200.ops.attention import packed_attention
    torch.manual_seed(3)
    qkv = torch.randn(2, 512, 3 * 4 * 64, device="cuda").to(torch.bfloat16)
    dout = torch.randn(2, 512, 4 * 64, device="cuda").to(torch.bfloat16)
    grads = []
    for _ in range(2):
        q = qkv.clone().requires_grad_(True)
        packed_attention(q, 4, True).backward(dout)
        grads.append(q.grad.clone())
    assert torch.equal(grads[0], grads[1])


# ------------------------------------------------------------------ the check scripts, under pytest
@pytest.mark.parametrize("script,marker,env", [
    ("scripts/gemm_check.py", "ALL_OK True", {}),
    ("scripts/gemm_check.py", "ALL_OK True", {"TDP_GEMM_EPI": "split", "TDP_GEMM_2CTA": "0"}),
    ("scripts/gemm2cta_check.py", "ALL_OK True", {}),
    ("scripts/fused_check.py", '"all_ok": true', {}),
    ("scripts/grouped_check.py", "ALL_OK True", {}),
    ("scripts/attn_check.py", "ALL_OK True", {}),
])
def test_kernel_check_scripts(script, marker, env):
    """The per-kernel numerics sweeps (all operand majors / ragged edges / every fused epilogue,
    grouped expert GEMM, attention) are part of the GPU gate, not only of the builder's runs."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, script)], cwd=ROOT, capture_output=True,
                       text=True, timeout=420, env={**os.environ, **env})
    assert r.returncode == 0 and marker in r.stdout, (script, r.stdout[-2500:], r.stderr[-2500:])
