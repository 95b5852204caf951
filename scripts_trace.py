"""Debug helper: analyse a VIPRS_B200_TRACE timeline (CTA 0 of the last sweep launch)."""
import sys
import numpy as np
NP_ = 4096
a = np.fromfile(sys.argv[1], dtype=np.uint64).astype(np.int64).reshape(10, 8, NP_)
t0 = a[a > 0].min()
def series(r, e):
    x = a[r, e]
    return {int(k): int(x[k] - t0) for k in np.nonzero(x)[0]}
PW0, PW1, P = series(9, 0), series(9, 1), series(9, 2)
BF0, BF1, BA, BC0, BC1 = series(0, 0), series(0, 1), series(0, 2), series(0, 4), series(0, 5)
B7A, B7C1 = series(7, 2), series(7, 5)
CW0, CW1, CS, CD = series(8, 0), series(8, 1), series(8, 2), series(8, 3)
npan = max(CD) + 1
print("panels", npan, "span cycles", max(CD.values()))
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (100, 110)
print("panel | prod: loopstart waitdone issue | bulk0: fullwait_start full_got A_done C_start C_done | b7 A_done C_done | chain: wait_start wait_done steps_done done")
for u in range(lo, hi):
    g = lambda d: d.get(u, -1)
    print(u, "|", g(PW0), g(PW1), g(P), "|", g(BF0), g(BF1), g(BA), g(BC0), g(BC1), "|", g(B7A), g(B7C1), "|", g(CW0), g(CW1), g(CS), g(CD))
def avg(x, y, lo, hi):
    ks = [k for k in x if k in y and lo <= k < hi]
    return np.mean([y[k] - x[k] for k in ks]) if ks else float('nan')
q = [0, npan // 4, npan // 2, 3 * npan // 4, npan]
print("phase quartiles (panels):", q)
for name, (x, y) in {"producer empty-wait": (PW0, PW1), "producer wait->issue": (PW1, P), "TMA issue->bulk0 full": (P, BF1),
                     "bulk0 full-wait": (BF0, BF1), "bulk0 A (incl window)": (BF1, BA), "bulk0 C": (BC0, BC1),
                     "A_done(b0)->chain go": (BA, CW1), "chain wait": (CW0, CW1), "chain steps": (CW1, CS), "chain epilogue": (CS, CD),
                     "chain done->C start(b0)": (CD, BC0)}.items():
    print(f"{name:26s}" + "".join(f"{avg(x, y, q[i], q[i + 1]):10.0f}" for i in range(4)))
print("chain period/panel        " + "".join(f"{(CD[q[i + 1] - 1] - CD[q[i]]) / (q[i + 1] - 1 - q[i]):10.0f}" for i in range(4)))
