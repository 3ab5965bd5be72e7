"""Weight-gradient GEMMs of the coupling backward (K = rows = 2^20, bf16 operands, fp32 result): which library call
is fastest for (rows x 256)^T (rows x 256) on matrices with row pitch 256 / 272."""
import torch, json


def timed(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows, U = 1 << 20, 256
out = {}
for pitch in (256, 272):
    ws = torch.randn(2, 4, rows, pitch, device="cuda", dtype=torch.bfloat16)
    h, d = ws[:, 0], ws[:, 3]
    out["mm_x2_pitch%d" % pitch] = timed(lambda: [torch.mm(h[n][:, :U].t(), d[n][:, :U], out_dtype=torch.float32) for n in range(2)])
    out["mm_x2_plus_sum_pitch%d" % pitch] = timed(lambda: [(torch.mm(h[n][:, :U].t(), d[n][:, :U], out_dtype=torch.float32), d[n][:, :U].sum(dim=0, dtype=torch.float32)) for n in range(2)])
    out["bmm_pitch%d" % pitch] = timed(lambda: torch.bmm(h[:, :, :U].transpose(1, 2), d[:, :, :U], out_dtype=torch.float32))
    if pitch > 256:
        out["mm_x2_ones264_pitch%d" % pitch] = timed(lambda: [torch.mm(h[n][:, :U + 8].t(), d[n][:, :U], out_dtype=torch.float32) for n in range(2)])
        out["bmm_ones264_pitch%d" % pitch] = timed(lambda: torch.bmm(h[:, :, :U + 8].transpose(1, 2), d[:, :, :U], out_dtype=torch.float32))
        out["mm_x2_ones272_pitch%d" % pitch] = timed(lambda: [torch.mm(h[n].t(), d[n][:, :U], out_dtype=torch.float32) for n in range(2)])
    out["sum_x2_pitch%d" % pitch] = timed(lambda: [d[n][:, :U].sum(dim=0, dtype=torch.float32) for n in range(2)])
    # the other direction: result^T = d^T h (cuBLAS sees other transposes)
    out["mmT_x2_pitch%d" % pitch] = timed(lambda: [torch.mm(d[n][:, :U].t(), h[n][:, :U], out_dtype=torch.float32) for n in range(2)])
    # split-K by hand: 8 chunks as a batched GEMM, summed
    def splitk(n):
        hh = h[n][:, :U].view(8, rows // 8, U) if pitch == U else None
        return hh
    del ws
ws = torch.randn(2, 4, rows, 272, device="cuda", dtype=torch.bfloat16)
x = torch.randn(rows, 40, device="cuda", dtype=torch.bfloat16)
d1, d3 = ws[:, 2], torch.randn(2, rows, 32, device="cuda", dtype=torch.bfloat16)
out["dW1_mm_x2"] = timed(lambda: [torch.mm(x.t(), d1[n][:, :U], out_dtype=torch.float32) for n in range(2)])
out["dW3_mm_x2_ones264"] = timed(lambda: [torch.mm(ws[n, 1][:, :U + 8].t(), d3[n], out_dtype=torch.float32) for n in range(2)])
out["dW3_mm_x2"] = timed(lambda: [torch.mm(ws[n, 1][:, :U].t(), d3[n], out_dtype=torch.float32) for n in range(2)])
print(json.dumps(out, indent=1))
