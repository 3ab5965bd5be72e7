"""Per-phase cycle trace of the tensor-core coupling kernel (CTA 0, lane 0 of each epilogue group).
Run on the GPU box:  python profiles/tc_phase_trace.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch_nf_b200 import _lib, ops
from torch_nf_b200.synthetic import synthetic_params

D, U, L, N = 64, 256, 2, 1 << 20
if len(sys.argv) > 2:
    N = int(sys.argv[2])
params = torch.tensor(synthetic_params([("RealNVP", L, U, True)], D, 1, seed=0)).cuda()
packed = ops.tc_pack(params[0], D, U, L, True)
z = torch.randn(1, N, D, device="cuda")
VARIANT = 0          # per-call diagnostic argument of tnf_coupling_tc (no process-global switches)
if len(sys.argv) > 1 and sys.argv[1] == "solo":
    VARIANT = 1 + 16
elif len(sys.argv) > 1 and sys.argv[1] == "pingpong":
    VARIANT = 1
elif len(sys.argv) > 1 and sys.argv[1].startswith("v"):
    VARIANT = int(sys.argv[1][1:])
for _ in range(3):
    ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, variant=VARIANT)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, variant=VARIANT)
e1.record(); torch.cuda.synchronize()
print("ms per launch: %.4f" % (e0.elapsed_time(e1) / 5))
dbg = torch.zeros(2048 * 4, dtype=torch.int64, device="cuda")
ops.coupling_tc(z, packed, D, U, L, True, ops.TNF_INVERSE, variant=VARIANT, debug=dbg)
torch.cuda.synchronize()
raw = dbg.cpu().numpy()
print("MMA warp (CTA 0): cycles waiting on epilogue %d, on weights %d, total %d (first layers %d, hidden layers %d, final layers %d)" % (raw[2040], raw[2041], raw[2042], raw[2043], raw[2044], raw[2045]))
d = raw[:2048].reshape(2, 512, 2)  # group g at int64 offset 1024*g
names = {602: "s read, acc freed", 603: "zin landed", 604: "z stored", 605: "pair synced", 250: "wait Hb net0", 260: "wait Hb net1", 350: "got  Hb net0", 360: "got  Hb net1", 100: "tile start", 101: "A1 published", 600: "tile end", 601: "y computed", 602: "next A1 published", 610: "A1 image written", 611: "A1 proxy fence done",
         700: "step: chunk landed", 701: "step: tanh+stores issued", 702: "step: chunk published", 703: "step: begin"}
for g in range(2):
    ev = [(int(t), int(c)) for t, c in d[g] if t != 0]
    if not ev:
        continue
    t0 = ev[0][1]
    print("group", g, "events", len(ev))
    prev = t0
    def label(t):
        return names.get(t) or ("wait H net%d l%d" % ((t - 200) // 10, (t - 200) % 10) if 200 <= t < 300 else
                                "got  H net%d l%d" % ((t - 300) // 10, (t - 300) % 10) if 300 <= t < 400 else
                                "wait F net%d" % (t - 400) if 400 <= t < 500 else "got  F net%d" % (t - 500))
    # mean duration of every phase over the steady tiles
    from collections import OrderedDict
    acc = OrderedDict()
    first = [i for i, (t, c) in enumerate(ev) if t == 100]
    if len(first) > 4:
        for i in range(first[2], first[-1]):
            key = (ev[i][0], ev[i + 1][0])
            acc.setdefault(key, []).append(ev[i + 1][1] - ev[i][1])
        print("  mean cycles per phase (steady tiles):")
        for (a_, b_), v in acc.items():
            print("    %-22s -> %-22s %7.0f  (n=%d)" % (label(a_), label(b_), np.mean(v), len(v)))
    if "-v" not in sys.argv:
        ev = []
    for i, (t, c) in enumerate(ev[60:100]):
        nm = names.get(t) or ("wait H net%d l%d" % ((t - 200) // 10, (t - 200) % 10) if 200 <= t < 300 else
                              "got  H net%d l%d" % ((t - 300) // 10, (t - 300) % 10) if 300 <= t < 400 else
                              "wait F net%d" % (t - 400) if 400 <= t < 500 else "got  F net%d" % (t - 500))
        print("  %8d  +%6d  %s" % (c - t0, c - prev, nm))
        prev = c
    # per-tile totals
    starts = [c for t, c in ev if t == 100]
    if len(starts) > 2:
        print("  cycles per tile (steady):", np.diff(starts)[1:].mean())

if "timeline" in sys.argv:
    # merged timeline of group 0, group 1 and the MMA warp (leader CTA 0) over two steady tiles
    ev = []
    for g in range(2):
        for tg, c in d[g]:
            if tg != 0:
                ev.append((int(c), "G%d %s" % (g, label(int(tg)))))
    m = raw[3072:3072 + 960].reshape(480, 2)
    jobs = {0: "t0", 1: "t1", 2: "tF", 3: "s0", 4: "s1", 5: "sF"}
    for tg, c in m:
        if tg != 0:
            tg = int(tg)
            kind = "issue begins" if tg < 2000 else "issued+committed"
            ev.append((int(c), "        MMA %s  group %d job %s" % (kind, (tg % 1000) // 100, jobs.get(tg % 100, str(tg % 100)))))
    ev.sort()
    starts = [c for c, s_ in ev if s_ == "G0 tile start"]
    if len(starts) > 12:
        lo, hi = starts[10], starts[12]
        prev = lo
        for c, s_ in ev:
            if lo <= c <= hi:
                print("%8d +%5d  %s" % (c - lo, c - prev, s_))
                prev = c
