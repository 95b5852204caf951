import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import viprs_b200 as vb
import oracle
from tests_util import make_block_ld
from test_round2_gpu import _hyper, _sweeps, _banded

def per_tile(got, ref, M):
    out = []
    den = np.max(np.abs(ref))
    for t0 in range(0, M, 2048):
        out.append(float(np.max(np.abs(got[t0:t0+2048] - ref[t0:t0+2048])) / den))
    return out

for blocks, U, T in (((6145,), np.float64, np.float64), ((8193,), np.float64, np.float64), ((10240,), np.float64, np.float64),
                     ((10240,), np.int8, np.float64)):
    rng = np.random.default_rng(55)
    P = make_block_ld(rng, blocks, U, T)
    hy = _hyper(rng, P["M"], T)
    for ns in (1, 2):
        ref = _sweeps(oracle.e_step, P, T, hy, ns)
        got = _sweeps(vb.cpp_e_step, P, T, hy, ns)
        print(blocks, U.__name__, "sweeps", ns, "eta per tile", ["%.1e" % v for v in per_tile(got["eta"], ref["eta"], P["M"])],
              "q", ["%.1e" % v for v in per_tile(got["q"], ref["q"], P["M"])], flush=True)

T = np.float32
rng = np.random.default_rng(77)
P = _banded(rng, 7001, 650, np.int8, T)
hy = _hyper(rng, P["M"], T)
P64 = dict(P, beta=P["beta"].astype(np.float64))
hy64 = tuple(np.asarray(a, np.float64) if isinstance(a, np.ndarray) else a for a in hy)
for ns in (1, 2, 3):
    ref = _sweeps(oracle.e_step, P, T, hy, ns)
    ref64 = _sweeps(oracle.e_step, P64, np.float64, hy64, ns)
    got = _sweeps(vb.cpp_e_step, P, T, hy, ns)
    from conftest import relmax
    print("banded i8 sweeps", ns, {k: ("%.1e/%.1e/%.1e" % (relmax(got[k], ref[k]), relmax(got[k], ref64[k]), relmax(ref[k], ref64[k]))) for k in ("eta", "var_gamma", "var_mu", "q")}, flush=True)
