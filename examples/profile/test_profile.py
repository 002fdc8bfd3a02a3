"""Module profiler demo (reference: examples/profile/test_profile.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn as nn
import torchdistpackage_b200 as tdp

class GEGLU(nn.Module):
    def __init__(self, d_in, d_out):
        super().__init__(); self.proj = nn.Linear(d_in, d_out * 2)
    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * torch.nn.functional.gelu(g)

dev = "cuda" if torch.cuda.is_available() else "cpu"
model = nn.Sequential(GEGLU(512, 2048), nn.Linear(2048, 512), nn.LayerNorm(512)).to(dev)
tdp.get_model_profile(model, args=(torch.randn(64, 128, 512, device=dev),), sort=False, max_depth=2)
