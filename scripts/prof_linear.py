import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zhusuan_b200 as zs
from zhusuan_b200 import fused as Fz
dev = "cuda"; R = 64 * 4096; K, J = 500, 784
h = torch.relu(torch.randn(R, K, device=dev)); W = torch.randn(J, K, device=dev) / K ** 0.5
b = torch.zeros(J, device=dev); x = (torch.rand(4096, J, device=dev) < 0.13).float()
g = torch.randn(R, device=dev)
wp, ws = Fz._tc_split(W); hp, hs = Fz._tc_split(h)
for _ in range(2):
    Fz._tc_linear(1, wp, ws, hp, hs, b, x, None, R, J, K)
    Fz._tc_linear(2, wp, ws, hp, hs, b, x, g, R, J, K)
torch.cuda.synchronize()
