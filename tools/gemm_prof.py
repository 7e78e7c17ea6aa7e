import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from gemm_check import gemm
n = 100000
g = torch.Generator(device="cuda").manual_seed(1)
S = torch.randn((n, 4160), device="cuda", generator=g); Bm = torch.randn((4160, 64), device="cuda", generator=g)
DM = torch.randn((n, 64), device="cuda", generator=g)
gemm(S, False, Bm, True, n, 64, 4160, engine=1)
gemm(S, True, DM, True, 4160, 64, n, splits=32, engine=1)
torch.cuda.synchronize()
