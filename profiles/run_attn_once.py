"""One Performer attention (5 tensor-core launches) + one conv module at the headline batch -- target of the ncu captures
of the control-network kernels (profiles/r02_ncu_control_*.txt)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddsp_b200 import core
from ddsp_b200.control import _orthogonal_gaussian_features
torch.manual_seed(0)
B, F, H = 64, 862, 8
x = torch.randn(B, F, 256, device='cuda')
w = torch.randn(1536, 256, device='cuda') / 16
hi, lo = core.split_tf32(w)
bias = torch.randn(1536, device='cuda') * 0.1
ps = (64 ** -0.25 * _orthogonal_gaussian_features(266, 64)).cuda().contiguous()
for _ in range(int(os.environ.get('REPS', '2'))):
    out = core.favor_attention(x, hi, lo, bias, ps, H)
    u = core.linear_ex(x, hi[:1024].contiguous(), weight_lo=lo[:1024].contiguous())
    s = core.glu_dwconv_silu(u, torch.randn(512, 31, device='cuda'), torch.randn(512, device='cuda'), u_bias=bias[:1024].contiguous())
torch.cuda.synchronize()
print('ok', out.abs().max().item())
