"""Debug helper: analyse a VIPRS_B200_TRACE timeline (CTA 0 of the last sweep launch)."""
import sys
import numpy as np
NP_ = 4096
a = np.fromfile(sys.argv[1], dtype=np.uint64).astype(np.int64).reshape(-1, 8, NP_)
t0 = a[a > 0].min()
fast = True
AW = range(4) if fast else range(8)
CW = range(4, 8) if fast else range(8)
def series(r, e):
    x = a[r, e]
    return {int(k): int(x[k] - t0) for k in np.nonzero(x)[0]}
def agg(ws, e, f):
    ss = [series(w, e) for w in ws]
    ks = set.intersection(*[set(s) for s in ss]) if ss else set()
    return {k: f(s[k] for s in ss) for k in ks}
PW0, PW1, P = series(9, 0), series(9, 1), series(9, 2)
AF1min, AF1max = agg(AW, 1, min), agg(AW, 1, max)
ADmin, ADmax = agg(AW, 2, min), agg(AW, 2, max)
C0min, C0max, C1min, C1max = agg(CW, 4, min), agg(CW, 4, max), agg(CW, 5, min), agg(CW, 5, max)
CW0, CW1, CS, CD = series(8, 0), series(8, 1), series(8, 2), series(8, 3)
npan = max(CD) + 1
print("kernel:", "fast" if fast else "generic", " panels", npan, " span cycles", max(CD.values()))
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (100, 106)
print("panel | prod: loopstart waitdone issue | A: full_got(min,max) done(min,max) | chain: wait_start go steps_done done | C: start(min,max) done(min,max)")
for u in range(lo, hi):
    g = lambda d: d.get(u, -1)
    print(u, "|", g(PW0), g(PW1), g(P), "|", g(AF1min), g(AF1max), g(ADmin), g(ADmax), "|", g(CW0), g(CW1), g(CS), g(CD), "|", g(C0min), g(C0max), g(C1min), g(C1max))
def avg(x, y, lo, hi):
    ks = [k for k in x if k in y and lo <= k < hi]
    return np.mean([y[k] - x[k] for k in ks]) if ks else float('nan')
q = [0, npan // 4, npan // 2, 3 * npan // 4, npan]
print("phase quartiles (panels):", q)
for name, (x, y) in {"producer empty-wait": (PW0, PW1), "producer wait->issue": (PW1, P), "TMA issue->first A got": (P, AF1min),
                     "A: got->done (slowest)": (AF1max, ADmax), "A spread (max-min done)": (ADmin, ADmax),
                     "A all done->chain go": (ADmax, CW1), "chain wait": (CW0, CW1), "chain steps": (CW1, CS), "chain epilogue": (CS, CD),
                     "chain done->C start(max)": (CD, C0max), "C: start->done (slowest)": (C0min, C1max),
                     "C done->next TMA issue": (C1max, {k - 4: v for k, v in P.items()})}.items():
    print(f"{name:26s}" + "".join(f"{avg(x, y, q[i], q[i + 1]):10.0f}" for i in range(4)))
A3min, A3max, A6min, A6max = agg(AW, 3, min), agg(AW, 3, max), agg(AW, 6, min), agg(AW, 6, max)
C6max = agg(CW, 6, max)
if A3max:
    for name, (x, y) in {"A: got->window done (max)": (AF1max, A3max), "A: window->groups done": (A3max, A6max), "A: epilogue": (A6max, ADmax),
                         "A fastest warp got->done": (AF1min, ADmin), "C: groups (slowest)": (C0min, C6max), "C: publish": (C6max, C1max),
                         "C fastest warp": (C0min, C1min)}.items():
        print(f"{name:26s}" + "".join(f"{avg(x, y, q[i], q[i + 1]):10.0f}" for i in range(4)))
def perwarp(w, e0, e1, lo, hi):
    x, y = series(w, e0), series(w, e1)
    return avg(x, y, lo, hi)
print("producer: top->prefetch issued->meta+shuffles(ev0)->gates(ev1)->meta stored(ev5)->tma issued(ev2)->released(ev6)->next top")
for a_, b_ in ((q[0], q[2]), (q[2], q[4])):
    P3, P4, P5, P6 = series(9, 3), series(9, 4), series(9, 5), series(9, 6)
    nxt = {k - 1: v for k, v in P3.items()}
    print("   ", " ".join(f"{avg(x, y, a_, b_):7.0f}" for x, y in ((P3, P4), (P4, PW0), (PW0, PW1), (PW1, P5), (P5, P), (P, P6), (P6, nxt))))
print("per-warp phase durations (first half | second half of the block)")
for w in AW:
    print(f"  A{w}: wait-full " + " | ".join(f"{perwarp(w, 0, 1, a, b):6.0f}" for a, b in ((q[0], q[2]), (q[2], q[4]))) +
          "  window " + " | ".join(f"{perwarp(w, 1, 3, a, b):6.0f}" for a, b in ((q[0], q[2]), (q[2], q[4]))) +
          "  groups " + " | ".join(f"{perwarp(w, 3, 6, a, b):6.0f}" for a, b in ((q[0], q[2]), (q[2], q[4]))) +
          "  g0 math " + " | ".join(f"{perwarp(w, 3, 4, a, b):6.0f}" for a, b in ((q[0], q[2]), (q[2], q[4]))) +
          "  g0 redux " + " | ".join(f"{perwarp(w, 4, 5, a, b):6.0f}" for a, b in ((q[0], q[2]), (q[2], q[4]))) +
          "  epilogue " + " | ".join(f"{perwarp(w, 6, 2, a, b):6.0f}" for a, b in ((q[0], q[2]), (q[2], q[4]))))
for w in CW:
    print(f"  C{w}: groups " + " | ".join(f"{perwarp(w, 4, 6, a, b):6.0f}" for a, b in ((q[0], q[2]), (q[2], q[4]))) +
          "  publish " + " | ".join(f"{perwarp(w, 6, 5, a, b):6.0f}" for a, b in ((q[0], q[2]), (q[2], q[4]))))
print("chain period/panel        " + "".join(f"{(CD[q[i + 1] - 1] - CD[q[i]]) / (q[i + 1] - 1 - q[i]):10.0f}" for i in range(4)))
# fine-grained A row-group trace (builds with -DVB_TRACE2): A warp index 3, second row group of every panel
T = [series(10, e) for e in range(8)]
if T[0]:
    steps = [("row-group top -> row bases loaded", 0, 1), ("-> top of the last tile's iteration", 1, 6), ("-> its addresses computed", 6, 7),
             ("-> its LDS.128 issued", 7, 2), ("-> its data landed", 2, 3), ("-> IDP.4A done", 3, 4), ("-> REDUX + selects done", 4, 5)]
    print("A (index 3) second row group, by block half (cycles):")
    for nm, a_, b_ in steps:
        print(f"  {nm:38s}" + " | ".join(f"{avg(T[a_], T[b_], a, b):7.0f}" for a, b in ((q[0], q[2]), (q[2], q[4]))))
