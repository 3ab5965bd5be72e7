"""Second probe of the weight-gradient GEMMs: operand order for the skinny products (dW1: 40 x 256, dW3: 272 x 32)."""
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


rows, U, DH = 1 << 20, 256, 32
ws = torch.randn(2, 4, rows, U + 16, device="cuda", dtype=torch.bfloat16)
d3 = torch.randn(2, rows, DH, device="cuda", dtype=torch.bfloat16)
out = {}
for xw in (40, 64, 128):
    xa = torch.randn(rows, xw, device="cuda", dtype=torch.bfloat16)
    out["dW1_xT_d_w%d" % xw] = timed(lambda: [torch.mm(xa.t(), ws[n, 2][:, :U], out_dtype=torch.float32) for n in range(2)])
    out["dW1_dT_x_w%d" % xw] = timed(lambda: [torch.mm(ws[n, 2][:, :U].t(), xa, out_dtype=torch.float32) for n in range(2)])
    out["dW1_dfullT_x_w%d" % xw] = timed(lambda: [torch.mm(ws[n, 2].t(), xa, out_dtype=torch.float32) for n in range(2)])
out["dW3_hT_d"] = timed(lambda: [torch.mm(ws[n, 1].t(), d3[n], out_dtype=torch.float32) for n in range(2)])
out["dW3_dT_h"] = timed(lambda: [torch.mm(d3[n].t(), ws[n, 1], out_dtype=torch.float32) for n in range(2)])
d3w = torch.randn(2, rows, 64, device="cuda", dtype=torch.bfloat16)
out["dW3_hT_d_w64"] = timed(lambda: [torch.mm(ws[n, 1].t(), d3w[n], out_dtype=torch.float32) for n in range(2)])
out["dW2_hT_d"] = timed(lambda: [torch.mm(ws[n, 0].t(), ws[n, 3][:, :U], out_dtype=torch.float32) for n in range(2)])
out["dW2_dT_h"] = timed(lambda: [torch.mm(ws[n, 3][:, :U].t(), ws[n, 0], out_dtype=torch.float32) for n in range(2)])
out["dW2_dfullT_h"] = timed(lambda: [torch.mm(ws[n, 3].t(), ws[n, 0], out_dtype=torch.float32) for n in range(2)])
# h1 and h2 side by side as ONE activation operand? (dW2 and dW3 have different deltas: not possible) - reference: pure read of one matrix
out["read_one_matrix_sum"] = timed(lambda: ws[0, 0].sum(dtype=torch.float32))
print(json.dumps(out, indent=1))
