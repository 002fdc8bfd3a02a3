from .gpt2 import GPT2, GPT2Config, GPT2Block, build_gpt2
from .moe_transformer import MoETransformer, MoEConfig, MoEBlock
