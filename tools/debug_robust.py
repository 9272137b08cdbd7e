import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, math
from oneprot_b200 import kernels as K
for n, d, amp in [(256, 512, 1.0), (300, 64, 2.0), (256, 64, 2.0), (300, 512, 1.0)]:
    g = torch.Generator().manual_seed(0)
    a = (amp * torch.randn(n, d, generator=g)).to(torch.bfloat16).cuda()
    b = (amp * torch.randn(n, d, generator=g)).to(torch.bfloat16).cuda()
    scale = torch.ones(1, device="cuda"); stats = torch.zeros(4, device="cuda")
    diag = torch.empty(n, device="cuda"); rs = torch.empty(n, device="cuda"); cs = torch.empty(n, device="cuda")
    K.rowstats(a, b, 0, diag, stats)
    K.fwd_sums(a, b, scale, stats, rs, cs)
    torch.cuda.synchronize()
    X = 1.4426950408889634 * (a.double() @ b.double().T)
    print(n, d, amp, "stats", stats.tolist(), "true xmax", X.max().item(), "rowsum min/max", rs.min().item(), rs.max().item(),
          "colsum min/max", cs.min().item(), cs.max().item(), "min rowmax", X.max(1).values.min().item(), "min colmax", X.max(0).values.min().item())
