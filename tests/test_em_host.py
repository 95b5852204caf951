"""
CPU tests of the scalar EM side (viprs_b200/em_host.py) and of the multi-GPU host logic (viprs_b200/parallel.py):
the M-step / ELBO / MSE computed from reduced sums must reproduce the numpy restatement of the reference's
VIPRS.m_step / elbo / mse (oracle/cpu.py, pinned against the goldens in tests/test_oracle.py), and the LD-block
sharding + one all-reduce must give every rank the whole-genome result (world_size 2, gloo).
"""
import os
import sys

import numpy as np
import pytest

from conftest import load_golden, ROOT


def _chroms(name):
    d, ch = load_golden(name)
    return d, ch


def _oracle_and_sums(oracle_built, name, mix=False, n_iter=3, theta=None, **kw):
    from oracle import cpu as ocpu
    from viprs_b200 import em_host
    d, ch = _chroms(name)
    keys = sorted(ch)
    ld = {c: (ch[c]["ld_data"], ch[c]["ld_indptr"], ch[c]["ld_left_bound"]) for c in keys}
    beta = {c: ch[c]["std_beta"] for c in keys}
    n = {c: ch[c]["n_per_snp"] for c in keys}
    cls = ocpu.OracleVIPRSMix if mix else ocpu.OracleVIPRS
    m = cls(ld, beta, n, **kw)
    m.initialize(dict(theta))
    return m, keys, em_host, ocpu


def test_slab_m_step_and_elbo_from_sums_match_numpy_restatement(oracle_built):
    theta = {"pi": 0.05, "sigma_epsilon": 0.7}
    m, keys, em_host, ocpu = _oracle_and_sums(oracle_built, "viprs_f32_f32.npz", theta=theta, float_precision="float32")
    hyp = em_host.SlabHyper(m.pi, m.sigma_epsilon, m.tau_beta, 0.0)
    seg_sizes = [m.shapes[c] for c in keys]
    for it in range(4):
        th = hyp.theta()
        m.e_step()
        S = np.stack([ocpu.sums_numpy(m.var_gamma[c], m.var_mu[c], m.eta[c], m.q[c], m.eta_diff[c], m.std_beta[c],
                                      m.n_per_snp[c], th) for c in keys])
        m.m_step()
        em_host.slab_m_step(S, seg_sizes, m.n_snps, hyp, False, False, False)
        assert np.isclose(hyp.pi[0], float(m.pi), rtol=2e-6)
        assert np.isclose(hyp.tau_beta[0], float(m.tau_beta), rtol=2e-6)
        assert np.isclose(hyp.sigma_g[0], float(m._sigma_g), rtol=2e-6)
        assert np.isclose(hyp.sigma_epsilon[0], float(m.sigma_epsilon), rtol=2e-6)
        assert np.isclose(em_host.slab_elbo(S, m.n, hyp, False)[0], float(m.elbo()), rtol=1e-6)
        assert np.isclose(em_host.slab_mse(S, hyp)[0], float(m.mse()), rtol=1e-5)
        assert np.isclose(em_host.max_eta_diff(S)[0], max(np.abs(v).max() for v in m.eta_diff.values()))
        # keep the two in lock-step (the oracle's float32 scalars differ from float64 ones at 1e-7)
        hyp.pi[0], hyp.tau_beta[0], hyp.sigma_epsilon[0] = float(m.pi), float(m.tau_beta), float(m.sigma_epsilon)


def test_fixed_sigma_epsilon_branch_of_elbo(oracle_built):
    theta = {"pi": 0.05, "sigma_epsilon": 0.7}
    m, keys, em_host, ocpu = _oracle_and_sums(oracle_built, "viprs_f32_f32.npz", theta=theta, float_precision="float64",
                                              fix_params={"sigma_epsilon": 0.7})
    hyp = em_host.SlabHyper(m.pi, m.sigma_epsilon, m.tau_beta, 0.0)
    th = hyp.theta()
    m.e_step()
    S = np.stack([ocpu.sums_numpy(m.var_gamma[c], m.var_mu[c], m.eta[c], m.q[c], m.eta_diff[c], m.std_beta[c],
                                  m.n_per_snp[c], th) for c in keys])
    m.m_step()
    em_host.slab_m_step(S, [m.shapes[c] for c in keys], m.n_snps, hyp, False, False, True)
    assert hyp.sigma_epsilon[0] == 0.7
    assert np.isclose(em_host.slab_elbo(S, m.n, hyp, True)[0], float(m.elbo()), rtol=1e-12)


def test_mixture_m_step_and_elbo_from_sums(oracle_built):
    d, _ = _chroms("viprsmix_f32_i16.npz")
    theta = {"pis": d["mix_pis"], "sigma_epsilon": 0.8}
    m, keys, em_host, ocpu = _oracle_and_sums(oracle_built, "viprsmix_f32_i16.npz", mix=True, theta=theta, K=4,
                                              float_precision="float32", dequantize_on_the_fly=True)
    hyp = em_host.MixHyper(m.pi, m.sigma_epsilon, m.tau_beta, m.d, 0.0)
    th0 = hyp.theta()
    for it in range(3):
        th = hyp.theta()
        m.e_step()
        S = np.stack([ocpu.sums_numpy(m.var_gamma[c], m.var_mu[c], m.eta[c], m.q[c], m.eta_diff[c], m.std_beta[c],
                                      m.n_per_snp[c], th, theta_logtau=th0, mixture=True) for c in keys])
        m.m_step()
        em_host.mix_m_step(S, m.n_snps, hyp, {})
        assert np.allclose(hyp.pi, np.asarray(m.pi, dtype=np.float64), rtol=5e-6)
        assert np.allclose(hyp.tau_beta, np.asarray(m.tau_beta, dtype=np.float64), rtol=5e-6)
        assert np.isclose(hyp.sigma_epsilon, float(m.sigma_epsilon), rtol=5e-6)
        assert np.isclose(em_host.mix_elbo(S, m.n, hyp, {}), float(m.elbo()), rtol=2e-6)
        assert np.isclose(em_host.mix_mse(S, hyp), float(m.mse()), rtol=1e-4)
        hyp.pi, hyp.tau_beta = np.asarray(m.pi, dtype=np.float64).copy(), np.asarray(m.tau_beta, dtype=np.float64).copy()
        hyp.sigma_epsilon = float(m.sigma_epsilon)


def test_find_blocks_and_partition():
    from viprs_b200 import parallel
    from tests_util import make_block_ld
    rng = np.random.default_rng(0)
    sizes = (257, 64, 1, 2, 33, 700, 17)
    for sym in (False, True):
        P = make_block_ld(rng, sizes, np.int8, np.float32, symmetric=sym)
        br = parallel.find_blocks(P["lb"], P["indptr"])
        assert list(np.diff(br)) == list(sizes)
    costs = parallel.block_costs(np.concatenate([[0], np.cumsum([4096] * 269)]))
    for world in (1, 2, 4, 8, 3):
        cut = parallel.partition_blocks(costs, world)
        n = np.diff(cut)
        assert cut[0] == 0 and cut[-1] == 269 and n.min() >= 269 // world - 1 and n.max() <= 269 // world + 2
    # more ranks than blocks: empty shards are allowed
    cut = parallel.partition_blocks(parallel.block_costs(np.array([0, 10, 30])), 4)
    assert cut[0] == 0 and cut[-1] == 2 and np.all(np.diff(cut) >= 0)


def _worker(rank, world, port, name, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cpu as ocpu
        from viprs_b200 import em_host, parallel
        from conftest import load_golden
        d, ch = load_golden(name)
        keys = sorted(ch)
        plan = parallel.shard_genome({c: ch[c] for c in keys}, world)[rank]
        local = {c: parallel.slice_chromosome(ch[c], *plan[c]) for c in keys}
        shapes = {c: len(ch[c]["std_beta"]) for c in keys}
        n_snps = sum(shapes.values())
        n_max = max(ch[c]["n_per_snp"].max() for c in keys)
        hyp = em_host.SlabHyper(0.05, np.float32(0.7), 0.05 * n_snps / (1 - np.float32(0.7)), 0.0)
        T = np.float32
        st = {c: {k: np.zeros(len(local[c]["std_beta"]), T) for k in ("var_mu", "eta", "q", "eta_diff")} for c in keys}
        for c in keys:
            st[c]["var_gamma"] = np.full(len(local[c]["std_beta"]), hyp.pi[0], T)
        ex = parallel.SumsExchange(len(keys), 1, ocpu.NSUMS, em_host.S_MAX_DIFF, rank, world, "cpu")
        hist = []
        for it in range(4):
            th = hyp.theta()
            S = np.zeros((len(keys), 1, ocpu.NSUMS))
            for i, c in enumerate(keys):
                L, s = local[c], st[c]
                if len(L["std_beta"]) == 0:
                    continue
                nn = L["n_per_snp"].astype(np.float64)
                vt = nn / th[0, 0] + th[0, 1]
                ul = (np.log(th[0, 2]) - np.log(1 - th[0, 2]) + .5 * (np.log(th[0, 1]) - np.log(vt))).astype(T)
                ocpu.e_step(L["ld_left_bound"], L["ld_indptr"], L["ld_data"], L["std_beta"].astype(T), s["var_gamma"],
                            s["var_mu"], s["eta"], s["q"], s["eta_diff"], ul, np.sqrt(.5 * vt).astype(T),
                            (nn / (vt * th[0, 0])).astype(T), 1.0, 1, True)
                S[i] = ocpu.sums_numpy(s["var_gamma"], s["var_mu"], s["eta"], s["q"], s["eta_diff"], L["std_beta"].astype(T),
                                       nn, th)
            if it % 2 == 0:
                G = ex.all_reduce(torch.from_numpy(S))
            else:                                   # the aliased path the device model uses: write into the send buffer
                ex.table().copy_(torch.from_numpy(S))
                G = ex.all_reduce(ex.table())
            em_host.slab_m_step(G, [shapes[c] for c in keys], n_snps, hyp, False, False, False)
            hist.append((float(em_host.slab_elbo(G, n_max, hyp, False)[0]), hyp.pi[0], hyp.tau_beta[0], hyp.sigma_epsilon[0],
                         float(em_host.max_eta_diff(G)[0])))
        q.put((rank, hist, {c: plan[c] for c in keys}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_em_over_gloo_matches_single_process(oracle_built, world):
    """LD-block sharding + one SUM all-reduce per iteration (the MAX slot rides one-hot) == whole-genome EM."""
    import torch.multiprocessing as mp
    name = "viprs_f32_f32.npz"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d, ch = load_golden(name)
    ref_elbo = d["em_hist_elbo"][:4]
    for rank, hist, plan in res:
        assert np.allclose([h[0] for h in hist], ref_elbo, rtol=1e-5), (rank, hist, ref_elbo)
        assert np.allclose([h[1] for h in hist], d["em_hist_pi"][:4], rtol=1e-4)
        assert np.allclose([h[3] for h in hist], d["em_hist_sigma_epsilon"][:4], rtol=1e-4)
        assert np.allclose([h[4] for h in hist], d["em_hist_max_eta_diff"][:4], rtol=1e-4)
    # all ranks agree bit-for-bit (same reduced table, same scalar code)
    assert all(r[1] == res[0][1] for r in res)
    # the shards tile every chromosome
    for c in res[0][2]:
        segs = sorted(r[2][c] for r in res if r[2][c][1] > r[2][c][0])
        assert segs[0][0] == 0 and segs[-1][1] == len(ch[c]["std_beta"])
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:]))


def test_fit_status_behaves_like_the_reference_optimize_result():
    """viprs_b200.optim restates viprs/utils/OptimizeResult.py; where the reference is importable (this container, not
    the GPU box) random update sequences must leave both in the same state."""
    import importlib.util
    import os
    ref_path = "/root/reference/viprs/utils/OptimizeResult.py"
    if not os.path.exists(ref_path):
        pytest.skip("reference tree not present")
    spec = importlib.util.spec_from_file_location("_ref_optres", ref_path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from viprs_b200.optim import FitStatus, Streak
    rng = np.random.default_rng(0)
    for trial in range(50):
        a, b = ref.OptimizeResult(), FitStatus()
        a.reset(); b.reset()
        ca, cb = ref.IterationConditionCounter(), Streak()
        f, it = 0.0, 0
        for step in range(40):
            f += rng.choice([-1.0, 1.0, 0.0, 0.5])
            stop = bool(rng.random() < 0.05)
            succ = bool(stop and rng.random() < 0.5)
            msg = rng.choice(["ok", "Maximum iterations reached", "diverged"])
            inc = bool(rng.random() < 0.9)
            a.update(f, stop_iteration=stop, success=succ, message=msg, increment=inc)
            b.update(f, stop_iteration=stop, success=succ, message=msg, increment=inc)
            it += int(rng.integers(1, 3))
            cond = bool(rng.random() < 0.7)
            ca.update(cond, it); cb.update(cond, it)
            assert (a.fun, a.nit, a.stop_iteration, a.success, a.message, a.error_on_termination, a.oscillation_counter,
                    bool(a.valid_optim_result)) == \
                   (b.fun, b.nit, b.stop_iteration, b.success, b.message, b.error_on_termination, b.oscillation_counter,
                    bool(b.valid_optim_result))
            assert ca.counter == cb.counter
