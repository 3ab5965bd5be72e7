import sys, time, torch
sys.path.insert(0, "/root/repo")
from torch_nf_b200 import ops
n = 1 << 20
z = torch.randn(1, n, 64, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
h = ops.to_host(z)
print("D2H to_host (pinned, 268MB): %.2f ms" % t(lambda: ops.to_host(z)))
print("D2H .cpu() pageable:         %.2f ms" % t(lambda: z.cpu()))
print("H2D pinned .cuda():          %.2f ms" % t(lambda: h.cuda()))
hp = z.cpu()
print("H2D pageable .cuda():        %.2f ms" % t(lambda: hp.cuda()))
print("is_pinned", h.is_pinned())
