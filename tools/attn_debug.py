import sys, os, math, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mobi_b200 import ops
rows, H, D, T, kern = [int(v) for v in sys.argv[1:6]]
torch.manual_seed(0)
q = (torch.randn(rows * H, T, D, device="cuda") * (D ** -0.5 * math.log2(math.e))).to(torch.bfloat16)
k = torch.randn(rows * H, T, D, device="cuda").to(torch.bfloat16)
v = torch.randn(rows * H, T, D, device="cuda").to(torch.bfloat16)
vt = v.transpose(1, 2).contiguous()
out = ops.attention(q, k, vt, rows, H, D, T, T, kernel=kern)
torch.cuda.synchronize()
ref = ops.attention(q, k, vt, rows, H, D, T, T, kernel=1)
torch.cuda.synchronize()
print("rows %d H %d D %d T %d kernel %d: ok, max diff vs one-tile kernel %.4g" % (rows, H, D, T, kern, (out.float() - ref.float()).abs().max().item()))
