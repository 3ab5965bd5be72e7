"""Discrete-event model of the two-group tensor-core coupling kernel (coupling_tc5.cu) used to rank scheduling changes
before writing them in CUDA.  Resources: ONE in-order MMA issuer + tensor pipe, ONE MUFU pipe shared by the two
epilogue groups (processor sharing: a lone tanh phase runs at ALONE of the MUFU rate, two concurrent ones at BOTH/2
each), fixed hand-off latencies.  Calibrated on profiles/r02_trace_tc5_*.txt (period ~25.5k cycles per two tiles)."""
import heapq, sys

ALONE, BOTH = 0.76, 0.97
L_COMMIT, L_ARRIVE = 150, 300      # MMA complete -> epilogue running ; epilogue done -> MMA warp sees it
GAP = 60                           # MMA warp bookkeeping between jobs
T_F0, T_HH, T_FIN = 320, 1088, 765 # tensor cycles: first layer, one N-half of the hidden layer, final layer
W_T1, W_T2H = 2048, 1024           # MUFU cycles of a tanh phase (T1) / of one half of T2
TAIL, RD = 3300, 120               # y compute + stores ; reading t or s out of TMEM


def simulate(order, n_tiles=40, opts=None):
    """order: list of (group, job) per period in MMA issue order; job in t0 t1a t1b tF s0 s1a s1b sF.
    opts: dict of scheduling features.  Returns the steady-state cycles per period (two tiles)."""
    opts = opts or {}
    tail = opts.get("tail", TAIL)
    early_k = opts.get("early_k", False)       # H_a's first K half may start when T1's first half is written
    fin_trail = opts.get("fin_trail", False)   # final job's first K half trails T2b
    # event times
    done = {}          # (g, tile, job) -> tensor completion time
    sig = {}           # (g, tile, name) -> time an epilogue signal becomes visible to the MMA warp
    # epilogue group state machines are simulated lazily with a global time-ordered loop
    mma_t = 0.0        # MMA warp free
    tens_t = 0.0       # tensor pipe free
    # We simulate by iterating "rounds": because of mutual dependencies use a fixed-point over time with an event queue.
    # Simpler: time-stepped co-simulation with small dt.
    dt = 4.0
    t = 0.0
    # group state: list of steps; each step = ("wait", key) | ("mufu", work, signal_after_half, signal_end) | ("busy", dur, signal) 
    def group_script(g):
        for k in range(n_tiles):
            for net in "ts":
                yield ("wait", (g, k, net + "0"))
                yield ("mufu", W_T1, (g, k, "e1h" + net), (g, k, "e1" + net))
                yield ("wait", (g, k, net + "1a"))
                yield ("mufu", W_T2H, None, (g, k, "e2a" + net))
                yield ("wait", (g, k, net + "1b"))
                yield ("mufu", W_T2H, None, (g, k, "e2" + net))
                yield ("wait", (g, k, net + "F"))
                yield ("busy", RD, (g, k, "e3" + net))
            yield ("busy", tail, None)
            yield ("mark", (g, k))
    scripts = [group_script(0), group_script(1)]
    cur = [next(scripts[0]), next(scripts[1])]
    rem = [None, None]       # remaining work of the current step
    half_sent = [False, False]
    marks = {0: [], 1: []}
    # MMA job list
    jobs = []
    for k in range(n_tiles + 1):
        for (g, j) in order:
            kk = k if (g, j) not in opts.get("prev_tile", ()) else k - 1
            if 0 <= kk < n_tiles:
                jobs.append((g, kk, j))
    ji = 0
    job_state = None   # (ready_to_issue_time)
    def dep(g, k, j):
        net = j[0]
        if j.endswith("0"):
            if net == "t":
                return [(g, k - 1, "e3s")] if k > 0 else []
            return [(g, k, "e3t")]
        if j.endswith("1a"):
            return [(g, k, ("e1h" if early_k else "e1") + net)]
        if j.endswith("1b"):
            return [(g, k, "e1" + net)]
        return [(g, k, ("e2a" if fin_trail else "e2") + net)]
    def dur(j):
        return T_F0 if j.endswith("0") else (T_FIN if j.endswith("F") else T_HH)
    pending_second = None
    while True:
        # MMA warp: issue next job if deps visible
        progressed = True
        while progressed and ji < len(jobs):
            progressed = False
            g, k, j = jobs[ji]
            ds = dep(g, k, j)
            if all(d in sig and sig[d] <= t for d in ds) and mma_t <= t:
                start = max(t, tens_t)
                d_ = dur(j)
                extra = 0.0
                if early_k and j.endswith("1a"):      # second K half needs the full T1: tensor waits for it
                    full = (g, k, "e1" + j[0])
                    # model: first half runs now, second half when e1 visible
                    first_end = start + d_ / 2
                    if full in sig:
                        second_start = max(first_end, sig[full])
                    else:
                        second_start = None
                    if second_start is None:
                        # cannot finish yet: block the MMA warp until e1 (in-order), emulate by waiting
                        break
                    end = second_start + d_ / 2
                elif fin_trail and j.endswith("F"):
                    full = (g, k, "e2" + j[0])
                    first_end = start + d_ / 2
                    if full not in sig:
                        break
                    end = max(first_end, sig[full]) + d_ / 2
                else:
                    end = start + d_
                tens_t = end
                done[(g, k, j)] = end + L_COMMIT
                mma_t = max(t, mma_t) + GAP + 0.5 * d_ * opts.get("issue_frac", 0.6)
                ji += 1
                progressed = True
        # epilogue groups
        active = [i for i in (0, 1) if cur[i] is not None and cur[i][0] == "mufu"]
        for i in (0, 1):
            st = cur[i]
            if st is None:
                continue
            if st[0] == "wait":
                if st[1] in done and done[st[1]] <= t:
                    cur[i] = next(scripts[i], None); rem[i] = None; half_sent[i] = False
            elif st[0] == "mufu":
                if rem[i] is None:
                    rem[i] = float(st[1])
                rate = ALONE if len(active) == 1 else BOTH / 2
                rem[i] -= rate * dt
                if st[2] is not None and not half_sent[i] and rem[i] <= st[1] / 2:
                    sig[st[2]] = t + L_ARRIVE; half_sent[i] = True
                if rem[i] <= 0:
                    if st[3] is not None:
                        sig[st[3]] = t + L_ARRIVE
                    cur[i] = next(scripts[i], None); rem[i] = None; half_sent[i] = False
            elif st[0] == "busy":
                if rem[i] is None:
                    rem[i] = float(st[1])
                rem[i] -= dt
                if rem[i] <= 0:
                    if st[2] is not None:
                        sig[st[2]] = t + L_ARRIVE
                    cur[i] = next(scripts[i], None); rem[i] = None
            elif st[0] == "mark":
                marks[i].append(t)
                cur[i] = next(scripts[i], None); rem[i] = None
        if cur[0] is None and cur[1] is None:
            break
        t += dt
        if t > 5e7:
            raise RuntimeError("stuck at job %s" % (jobs[ji],) if ji < len(jobs) else "stuck")
    m = marks[0]
    return (m[-1] - m[len(m) // 2]) / (len(m) - 1 - len(m) // 2)


def order_shift(sh, halves_interleaved=False):
    seq = ["t0", "t1a", "t1b", "tF", "s0", "s1a", "s1b", "sF"]
    jobs6 = [["t0"], ["t1a", "t1b"], ["tF"], ["s0"], ["s1a", "s1b"], ["sF"]]
    out, prev = [], set()
    for s in range(6):
        out += [(0, j) for j in jobs6[s]]
        j1 = (s - sh) % 6
        for j in jobs6[j1]:
            out.append((1, j))
            if s < sh:
                prev.add((1, j))
    return out, prev


if __name__ == "__main__":
    for sh in (1, 2, 3):
        order, prev = order_shift(sh)
        base = simulate(order, opts={"prev_tile": prev})
        print("shift %d: period %.0f cycles per two tiles" % (sh, base))
    order, prev = order_shift(2)
    for name, o in (("baseline", {}), ("tail 1500", {"tail": 1500}), ("tail 500", {"tail": 500}), ("early K half", {"early_k": True}),
                    ("final trails T2b", {"fin_trail": True}), ("early K + fin trail", {"early_k": True, "fin_trail": True}),
                    ("early K + fin trail + tail 1500", {"early_k": True, "fin_trail": True, "tail": 1500})):
        o = dict(o); o["prev_tile"] = prev
        print("%-34s %.0f" % (name, simulate(order, opts=o)))
