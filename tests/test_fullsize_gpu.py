"""The bench batch itself (BASELINE configs[1]/[3] shape: 16 384 registrations of an Oxford-shape scan against a 10-scan submap, ~2.9 M cell
pairs, tables larger than L2), built by the product path exactly as bench.py builds it, checked

  * against the CPU oracle over ALL of its pairs (one Jet<4> pass takes the oracle about a second on one thread), loss and no loss;
  * through size-independent properties: the per-pair EMIT output folded on the host reproduces the FUSED records, the association
    obeys its structural invariants everywhere and equals the oracle's on a sample of problems, the batched solver leaves every
    registration converged with a cost no higher than where it started.
"""
import numpy as np
import pytest

from randt_slam_b200 import capi, params as P

pytestmark = pytest.mark.gpu

N_PROBLEMS = 16384


@pytest.fixture(scope="module")
def batch(gpu_ctx):
    import bench
    prob, poses, st, host = bench.build_problem(gpu_ctx, capi, P.OXFORD, N_PROBLEMS, 1)
    yield dict(prob=prob, poses=poses, st=st, host=host)
    prob.close()


def _seg_sum(v, seg):
    """sum of v over [seg[s], seg[s+1]) for every s, empty segments -> 0"""
    c = np.concatenate([np.zeros((1,) + v.shape[1:]), np.cumsum(v, axis=0)])
    return c[seg[1:]] - c[seg[:-1]]


def test_batch_has_the_bench_shape(batch):
    st = batch["st"]
    assert st["segments"] == N_PROBLEMS and st["pairs"] > 2_500_000
    assert 48 * (st["n_m"] + st["n_f"]) > 126e6          # tables alone exceed the L2


@pytest.mark.parametrize("with_loss", [False, True])
def test_fused_equals_the_oracle_over_the_whole_batch(oracle, batch, with_loss):
    p = P.OXFORD
    prob, poses, host = batch["prob"], batch["poses"], batch["host"]
    loss = capi.make_loss(capi.LOSS_BARRON, p.loop_closure_scale, p.loss_function_convexity, 1.0, 1.0) if with_loss else None
    lt = (capi.LOSS_BARRON, p.loop_closure_scale, p.loss_function_convexity, 1.0, 1.0) if with_loss else (capi.LOSS_NONE, 1.0, -2.0, 1.0, 1.0)
    got = prob.eval_fused(poses, loss)
    want, _ = oracle.fused_batch(0, host["cells_m"], host["cells_f"], host["pm"], host["pf"], host["seg"], poses, lt, n_threads=oracle.hw_threads())
    H = got[:, :16].reshape(-1, 4, 4)
    scale_H = np.abs(want["H"]).max(axis=(1, 2), keepdims=True) + 1e-300
    scale_g = np.abs(want["g"]).max(axis=1, keepdims=True) + 1e-300
    assert np.max(np.abs(H - want["H"]) / scale_H) < 1e-9
    assert np.max(np.abs(got[:, 16:20] - want["g"]) / scale_g) < 1e-9
    assert np.allclose(got[:, capi.FUSED_COST], want["cost"], rtol=1e-11, atol=0)
    assert np.allclose(got[:, capi.FUSED_MAXR], want["max_r"], rtol=1e-12, atol=0)
    assert np.array_equal(got[:, capi.FUSED_N], want["n"])
    assert np.array_equal(got[:, capi.FUSED_N], np.diff(host["seg"]).astype(np.float64))


def test_emit_folded_on_the_host_reproduces_fused(batch):
    """sum_p J_p^T J_p, sum_p J_p^T r_p, 1/2 sum r_p^2, max r_p per registration from the per-pair output == the fused records (no loss)"""
    prob, poses, seg = batch["prob"], batch["poses"], batch["host"]["seg"].astype(np.int64)
    r, J = prob.eval_emit(poses)
    fused = prob.eval_fused(poses, None)
    assert np.isfinite(r).all() and np.isfinite(J).all() and (r >= 0).all()
    H = _seg_sum((J[:, :, None] * J[:, None, :]).reshape(-1, 16), seg)
    g = _seg_sum(J * r[:, None], seg)
    cost = 0.5 * _seg_sum((r * r)[:, None], seg)[:, 0]
    scale = np.abs(H).max(axis=1, keepdims=True) + 1e-300
    assert np.max(np.abs(fused[:, :16] - H) / scale) < 1e-9          # cumulative-sum differences on the host side set this bound
    assert np.max(np.abs(fused[:, 16:20] - g) / (np.abs(g).max(axis=1, keepdims=True) + 1e-300)) < 1e-8
    assert np.allclose(fused[:, capi.FUSED_COST], cost, rtol=1e-9, atol=1e-12)
    nz = np.flatnonzero(np.diff(seg) > 0)
    assert np.array_equal(fused[nz, capi.FUSED_MAXR], np.maximum.reduceat(r, seg[:-1][nz]))   # a max is order-independent: exact


def test_association_invariants_and_oracle_sample(oracle, batch):
    p = P.OXFORD
    k = p.n_results_nn_lookup
    host, poses = batch["host"], batch["poses"]
    pm, pf, seg = host["pm"].astype(np.int64), host["pf"].astype(np.int64), host["seg"].astype(np.int64)
    m_off, f_off = host["m_off"].astype(np.int64), host["f_off"].astype(np.int64)
    owner = np.repeat(np.arange(N_PROBLEMS), np.diff(seg))
    assert (pm >= m_off[owner]).all() and (pm < m_off[owner + 1]).all()       # a pair never leaves its registration
    assert (pf >= f_off[owner]).all() and (pf < f_off[owner + 1]).all()
    assert (np.diff(pm) >= 0).all()                                           # moving cells in order (one residual block after another)
    assert np.bincount(pm, minlength=int(m_off[-1])).max() <= k               # at most k neighbours per moving cell
    same = pm[1:] == pm[:-1]
    assert (pf[1:][same] != pf[:-1][same]).all()                              # and never the same neighbour twice in a row
    for b in np.random.default_rng(0).choice(N_PROBLEMS, 24, replace=False):
        im, jf = oracle.associate(host["cells_f"][f_off[b]:f_off[b + 1]], host["slot"][b], p.size_x, p.size_y, p.resolution,
                                  p.max_neighbor_linf_distance, host["cells_m"][m_off[b]:m_off[b + 1]], poses[b], k)
        a, z = seg[b], seg[b + 1]
        assert np.array_equal(pm[a:z] - m_off[b], im) and np.array_equal(pf[a:z] - f_off[b], jf)


def test_every_registration_of_the_batch_converges(batch):
    p = P.OXFORD
    prob, poses = batch["prob"], batch["poses"]
    loss = capi.make_loss(capi.LOSS_BARRON, p.loop_closure_scale, p.loss_function_convexity, 1.0, 1.0)
    opt = capi.solver_options(use_manifold=1, gnc_loss_scale=p.loss_function_scale, gnc_divisor=p.gnc_control_parameter_divisor,
                              gnc_max_steps=p.loop_closure_gnc_steps, max_num_iterations=p.max_iteration)
    start = prob.eval_fused(poses, loss)[:, capi.FUSED_COST]
    out, res = prob.register_batch(poses, loss, opt)
    assert (res[:, capi.REG_STATUS] == 0).all()
    assert (res[:, capi.REG_TERMINATION] == 0).mean() > 0.99                   # CONVERGENCE (a few may stop on the iteration limit)
    assert np.allclose(np.hypot(out[:, 0], out[:, 1]), 1.0, atol=1e-12)        # the manifold keeps (cos, sin) on the unit circle
    end = prob.eval_fused(out, loss)[:, capi.FUSED_COST]                       # mu = 1: the last GNC stage's objective
    assert (end <= start * (1 + 1e-9)).mean() > 0.99 and np.median(end / start) < 0.9
    assert np.allclose(end, res[:, capi.REG_FINAL_COST], rtol=1e-3, atol=1e-9)  # final_cost: lowest cost the last solve has seen
    again, res2 = prob.register_batch(poses, loss, opt)                        # bitwise reproducible
    assert np.array_equal(out, again) and np.array_equal(res, res2)
