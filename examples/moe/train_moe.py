"""MoE transformer with expert parallel x replicated-expert data parallel
(reference recipe: ddp/moe_dp.md; the MoE layer itself is new -- the reference delegates it to
DeepSpeed/FastMoE forks, explore/moe/ds_fmoe_main.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from _common import init, log, tdp
from torchdistpackage_b200.models import MoETransformer, MoEConfig

rank, world, dev = init(__doc__)
tdp.tpc.setup_process_groups([("data", world)])
ep = 2 if world % 2 == 0 else 1
tdp.tpc.build_moe_groups(moe_ep_size=ep)
cfg = MoEConfig.tiny()
tdp.fix_rand(0)
model = MoETransformer(cfg, ep_group=tdp.tpc.get_group("moe_ep")).to(dev)
model = model.to(torch.bfloat16 if dev.type == "cuda" else torch.float32)
model._ddp_params_and_buffers_to_ignore = model.ddp_ignore_names()
tdp.create_moe_dp_hooks(model.expert_parameters(), tdp.tpc.get_group("moe_dp"),
                        tdp.tpc.get_ranks_in_group("moe_dp")[0])
ddp = tdp.NaiveDDP(model, process_group=tdp.tpc.get_group("data"), gradient_as_bucket_view=True)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
tok = torch.randint(0, cfg.vocab_size, (4, cfg.seq_len + 1), device=dev)
for it in range(5):
    opt.zero_grad(set_to_none=False)
    loss = ddp(tok[:, :-1], tok[:, 1:]); loss.backward()
    ddp.reduce_gradients(); tdp.moe_dp_iter_step()
    opt.step()
    log(rank, f"step {it} loss {loss.item():.4f}")
