"""Launch target for ncu: one mask forward + backward at the bench shape (fp32), three times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdqe_cvpr2023_b200 import ops
B, Q, K = 1, 196, 32
coeff = torch.tanh(torch.randn(B, Q, K, device="cuda")); proto = torch.randn(B, K, 4, 96, 160, device="cuda"); go = torch.randn(B, Q, 4, 96, 160, device="cuda")
for _ in range(3):
    ops.mask_logits_forward(coeff, proto)
    ops.mask_logits_backward(coeff, proto, go)
torch.cuda.synchronize()
