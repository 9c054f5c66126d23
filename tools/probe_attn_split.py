import sys, torch
sys.path.insert(0, '/root/repo')
from more4d_b200 import _lib, ops
from oracle import dit_oracle as O
torch.set_grad_enabled(False)
def rnd(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16)
ar = O.Arith(True)
print("start", flush=True)
_lib.lib().m4d_set_debug_flags(0x6000000)
for (B, Lq, Lk, N) in [(1, 128, 128, 1), (1, 256, 384, 2), (2, 300, 257, 2), (1, 1000, 3000, 3)]:
    q, k, v = rnd((B, Lq, N, 128), 1, 2.0), rnd((B, Lk, N, 128), 2, 2.0), rnd((B, Lk, N, 128), 3)
    out = ops.attention(q.cuda(), k.cuda(), v.cuda())
    torch.cuda.synchronize()
    ref = O.attention(q, k, v, None, ar)
    e = float((out.float().cpu() - ref).norm() / ref.norm())
    print(f"split B{B} Lq{Lq} Lk{Lk} N{N}: rel={e:.3e}", flush=True)
