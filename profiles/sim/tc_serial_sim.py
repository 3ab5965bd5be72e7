"""Model of the 'all 16 epilogue warps serve both groups in a fixed global phase order' variant: the MUFU-bound
phases are strictly serialised (each runs on 4 warps per sub-partition at EFF of the MUFU rate); the MMA issuer works
through its static job list; a phase waits for its MMA job (commit + L_COMMIT), a job waits for the phase that frees
its accumulator / writes its operand (+ L_ARRIVE)."""
EFF = 0.95
L_COMMIT, L_ARRIVE, GAP = 150, 300, 60
T_F0, T_HH, T_FIN = 320, 1088, 765
W_T1, W_T2H = 2048, 1024
RD, TAIL = 120, 1650      # tail executed by all 16 warps: half of the 8-warp tail


def simulate(phase_order, job_order, n_tiles=30, tail=TAIL, rd=RD, eff=EFF):
    ph_done, job_done = {}, {}
    phases = [(g, p, k + off) for k in range(n_tiles + 2) for (g, p, off) in phase_order if 0 <= k + off < n_tiles]
    jobs = [(g, j, k + off) for k in range(n_tiles + 2) for (g, j, off) in job_order if 0 <= k + off < n_tiles]
    need_job = {"T1t": "t0", "T2t": "t1a", "RDt": "tF", "T1s": "s0", "T2s": "s1a", "TAIL": "sF"}
    need_phase = {"t0": ("TAIL", -1), "t1": ("T1t", 0), "tF": ("T2t", 0), "s0": ("RDt", 0), "s1": ("T1s", 0), "sF": ("T2s", 0)}
    pi = ji = 0
    t_epi = t_mma = t_tens = 0.0
    marks = []
    while pi < len(phases):
        progressed = False
        while ji < len(jobs):
            g, j, k = jobs[ji]
            p, off = need_phase[j]
            key = (g, k + off, p)
            if k + off >= 0 and key not in ph_done:
                break
            ready = ph_done.get(key, 0.0)
            start = max(t_mma, ready)
            if j in ("t1", "s1"):
                s0 = max(start, t_tens)
                job_done[(g, k, j + "a")] = s0 + T_HH + L_COMMIT
                job_done[(g, k, j + "b")] = s0 + 2 * T_HH + L_COMMIT
                t_tens = s0 + 2 * T_HH
                t_mma = start + GAP + 0.6 * T_HH
            else:
                d = T_F0 if j.endswith("0") else T_FIN
                s0 = max(start, t_tens)
                job_done[(g, k, j)] = s0 + d + L_COMMIT
                t_tens = s0 + d
                t_mma = start + GAP + 0.5 * d
            ji += 1
            progressed = True
        g, p, k = phases[pi]
        key = (g, k, need_job[p])
        if key in job_done:
            start = max(t_epi, job_done[key])
            if p in ("T1t", "T1s"):
                end = start + W_T1 / eff
            elif p in ("T2t", "T2s"):
                mid = start + W_T2H / eff
                end = max(mid, job_done[(g, k, need_job[p][:-1] + "b")]) + W_T2H / eff
            elif p == "RDt":
                end = start + rd
            else:
                end = start + tail
            ph_done[(g, k, p)] = (start + rd if p == "TAIL" else end) + L_ARRIVE
            t_epi = end
            if g == 0 and p == "TAIL":
                marks.append(end)
            pi += 1
            progressed = True
        if not progressed:
            raise RuntimeError("deadlock at phase %s job %s" % (phases[pi], jobs[ji] if ji < len(jobs) else None))
    m = marks
    return (m[-1] - m[len(m) // 2]) / (len(m) - 1 - len(m) // 2)


if __name__ == "__main__":
    P = ["T1t", "T2t", "RDt", "T1s", "T2s", "TAIL"]
    J = ["t0", "t1", "tF", "s0", "s1", "sF"]
    for sh in range(1, 6):
        po, jo = [], []
        for i in range(6):
            po.append((0, P[i], 0))
            po.append((1, P[(i - sh) % 6], 0 if i >= sh else -1))
        for i in range(6):
            jo.append((0, J[i], 0))
            jo.append((1, J[(i - sh) % 6], 0 if i >= sh else -1))
        for jshift in (0, 1, 2):
            jo2 = jo[jshift:] + [(g, j, o + 1) for (g, j, o) in jo[:jshift]]
            for eff in (0.95, 0.90):
                try:
                    r = simulate(po, jo2, eff=eff)
                    print("phase shift %d, job list rotated %d, eff %.2f: period %.0f" % (sh, jshift, eff, r))
                except RuntimeError as e:
                    print("phase shift %d, job list rotated %d: %s" % (sh, jshift, e))
