"""Which physical layouts does the library attention use for O / dQ / dK / dV when q, k, v are
strided views of a packed qkv tensor?  (decides how many layout copies packed_attention needs)"""
import torch
import torch.nn.functional as F
B, T, H, dh = 16, 1024, 12, 64
qkv = torch.randn(B, T, 3, H, dh, device="cuda", dtype=torch.bfloat16)
q, k, v = (qkv[:, :, i].transpose(1, 2).detach().requires_grad_(True) for i in range(3))
o = F.scaled_dot_product_attention(q, k, v, is_causal=True)
print("q", tuple(q.shape), q.stride()); print("o", tuple(o.shape), o.stride(), "o^T contiguous:", o.transpose(1, 2).is_contiguous())
dout = torch.randn(B, T, H * dh, device="cuda", dtype=torch.bfloat16)
dview = dout.reshape(B, T, H, dh).transpose(1, 2)
print("dO view strides", dview.stride(), "== o.stride():", dview.stride() == o.stride())
dq, dk, dv = torch.autograd.grad(o, (q, k, v), dview)
for n, g in (("dq", dq), ("dk", dk), ("dv", dv)):
    print(n, g.stride(), "data_ptr", g.data_ptr(), "storage_offset", g.storage_offset())
