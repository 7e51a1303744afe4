"""K3 parity (GPU vs CPU oracle): residuals, Jacobians, fused normal equations, loss variants, sweep.

Tolerance: north_star asks for 1e-5 relative to the reference's Eigen/Ceres path; the fp64 closed form agrees with the
oracle's dual-number path to ~1e-12, asserted here at 1e-9 (row-wise relative for J, see rowwise()).
"""
import math

import numpy as np
import pytest

from randt_slam_b200 import capi, params as P, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL = 1e-9


def rowwise(Jg, Jo):
    scale = np.maximum(np.max(np.abs(Jo), axis=1, keepdims=True), 1e-30)
    return float(np.max(np.abs(Jg - Jo) / scale))


def loss_tuple(l):
    return (l.kind, l.scale, l.alpha, l.mu, l.weight)


@pytest.mark.parametrize("preset", ["oxford", "c1", "indoor"])
def test_emit_matches_oracle_on_scans(oracle, gpu_ctx, preset):
    p = {"oxford": P.OXFORD, "c1": P.C1, "indoor": P.INDOOR}[preset]
    case = H.make_registration_case(oracle, p, seed=3)
    im, jf = case["im"], case["jf"]
    assert len(im) > 50
    pose = case["pose0"].copy(); pose[:2] *= 1.0007   # un-normalised complex part: exercises the ambient derivative
    prob = gpu_ctx.problem_create(case["moving"]["cells"], case["fixed"]["cells"], im, jf, [0, len(im)])
    r, J = prob.eval_emit(pose)
    ro, Jo = oracle.eval_pairs(0, case["moving"]["cells"], case["fixed"]["cells"], im, jf, pose, 0)
    assert H.rel_err(r, ro) < TOL
    assert np.max(np.abs(r - ro) / ro) < TOL
    assert rowwise(J, Jo) < 1e-8
    r2, J2 = prob.eval_emit(pose, want_jac=False)
    assert J2 is None and np.array_equal(r, r2)
    assert gpu_ctx.take_bad_pairs() == 0


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_all_variants_random_cells(oracle, gpu_ctx, variant):
    rng = np.random.default_rng(10 + variant)
    cm = H.random_cells(rng, 400); cf = H.random_cells(rng, 900)
    # make every pair geometrically close so residuals are O(1)
    im = rng.integers(0, 400, 3000).astype(np.uint32); jf = rng.integers(0, 900, 3000).astype(np.uint32)
    cf2 = cf.copy()
    seg = [0, 700, 700, 1900, 3000]   # includes an empty segment
    if variant <= 1:
        poses = np.stack([synth.pose_to_se2(*rng.uniform(-1, 1, 3) * [1, 1, 0.5]) * [1.001, 1.001, 1, 1] for _ in range(4)])
    else:
        poses = rng.uniform(-1, 1, (4, 3)) * [1, 1, 4.0]
    prob = gpu_ctx.problem_create(cm, cf2, im, jf, seg)
    r, J = prob.eval_emit(poses, variant=variant)
    for s in range(4):
        a, b = seg[s], seg[s + 1]
        ro, Jo = oracle.eval_pairs(variant, cm, cf2, im[a:b], jf[a:b], poses[s], 0)
        if b > a:
            assert np.max(np.abs(r[a:b] - ro) / ro) < TOL
            assert rowwise(J[a:b], Jo) < 1e-8


LOSSES = [
    capi.make_loss(capi.LOSS_NONE),
    capi.make_loss(capi.LOSS_BARRON, 1.0, -2.0, 1.0, 5000.0 / 300),
    capi.make_loss(capi.LOSS_BARRON, 2.0, -1.0, 3.3, 1.0),
    capi.make_loss(capi.LOSS_BARRON, 2.0, -1.5, 1.21, 0.7),
    capi.make_loss(capi.LOSS_BARRON, 1.5, 0.03, 1.0, 1.0),
    capi.make_loss(capi.LOSS_BARRON, 1.5, 1.0, 2.0, 1.0),
    capi.make_loss(capi.LOSS_BARRON, 1.5, 2.5, 2.0, 1.0),
    capi.make_loss(capi.LOSS_WELSCH, 1.5, 0.0, 2.0, 1.3),
]


@pytest.mark.parametrize("li", range(len(LOSSES)))
@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_fused_matches_oracle(oracle, gpu_ctx, li, variant):
    loss = LOSSES[li]
    case = H.make_registration_case(oracle, P.OXFORD, seed=5)
    cm, cf, im, jf = case["moving"]["cells"], case["fixed"]["cells"], case["im"], case["jf"]
    n = len(im)
    seg = [0, n // 3, n]
    if variant <= 1:
        poses = np.stack([case["pose0"], synth.pose_to_se2(0.55, -0.35, 0.025) * [0.93, 0.93, 1, 1]])
    else:
        poses = np.array([[0.5, -0.3, 0.02], [0.55, -0.35, 0.025 + 2 * math.pi]])
    prob = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    for want_jac in (True, False):
        out = capi.unpack_fused(prob.eval_fused(poses, loss, want_jac=want_jac, variant=variant))
        for s in range(2):
            a, b = seg[s], seg[s + 1]
            fo = oracle.fused(variant, cm, cf, im[a:b], jf[a:b], poses[s], loss_tuple(loss), want_jac)
            assert abs(out["cost"][s] - fo["cost"]) <= 1e-9 * abs(fo["cost"])
            assert abs(out["max_r"][s] - fo["max_r"]) <= 1e-10 * fo["max_r"]
            assert abs(out["sum_sq"][s] - fo["sum_sq"]) <= 1e-9 * fo["sum_sq"]
            assert out["n"][s] == b - a
            if want_jac:
                assert H.rel_err(out["H"][s], fo["H"]) < 1e-8
                assert H.rel_err(out["g"][s], fo["g"]) < 1e-8
            else:
                assert np.all(out["H"][s] == 0) and np.all(out["g"][s] == 0)


def test_fused_per_segment_mu_and_multi_tile(oracle, gpu_ctx):
    """segments larger than one 512-pair tile go through the last-CTA fold; mu may differ per segment."""
    rng = np.random.default_rng(77)
    cm = H.random_cells(rng, 1500, extent=5.0); cf = H.random_cells(rng, 2500, extent=5.0)
    P_ = 9000
    im = np.sort(rng.integers(0, 1500, P_)).astype(np.uint32); jf = rng.integers(0, 2500, P_).astype(np.uint32)
    seg = [0, 100, 5000, 5001, 9000]
    poses = np.stack([synth.pose_to_se2(*rng.uniform(-0.2, 0.2, 3)) for _ in range(4)])
    mus = np.array([1.0, 2.5, 7.0, 1.3])
    loss = capi.make_loss(capi.LOSS_BARRON, 1.0, -2.0, 1.0, 0.01)
    prob = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    out1 = prob.eval_fused(poses, loss, mu_per_seg=mus)
    out2 = prob.eval_fused(poses, loss, mu_per_seg=mus)
    assert np.array_equal(out1, out2), "fused reduction must be deterministic run to run"
    o = capi.unpack_fused(out1)
    fo, _ = oracle.fused_batch(0, cm, cf, im, jf, seg, poses, (loss.kind, loss.scale, loss.alpha, 1.0, loss.weight), mu_per_seg=mus)
    assert H.rel_err(o["H"], fo["H"]) < 1e-9 and H.rel_err(o["g"], fo["g"]) < 1e-9
    assert np.allclose(o["cost"], fo["cost"], rtol=1e-10) and np.allclose(o["max_r"], fo["max_r"], rtol=1e-12)
    assert np.array_equal(o["n"], fo["n"])


def test_sweep_costs(oracle, gpu_ctx):
    case = H.make_registration_case(oracle, P.OXFORD, seed=6)
    cm, cf, im, jf = case["moving"]["cells"], case["fixed"]["cells"], case["im"], case["jf"]
    rng = np.random.default_rng(1)
    poses = np.stack([synth.pose_to_se2(*(np.array([0.6, -0.4, 0.03]) + rng.uniform(-1, 1, 3) * [2.0, 2.0, 0.2])) for _ in range(333)])
    loss = capi.make_loss(capi.LOSS_BARRON, 0.5, -2.0, 1.0, 1.0)
    prob = gpu_ctx.problem_create(cm, cf, im, jf, [0, len(im)])
    c = prob.sweep_costs(0, poses, loss)
    co = oracle.sweep_costs(0, cm, cf, im, jf, loss_tuple(loss), poses)
    assert np.max(np.abs(c - co) / co) < 1e-10


def test_degenerate_pairs_are_flagged(oracle, gpu_ctx):
    cm = np.zeros((2, 12), np.float32); cf = np.zeros((2, 12), np.float32)
    cm[0, :3] = [1, 2, 90]; cm[0, 3:] = np.eye(3).reshape(9)        # fine
    cf[0, :3] = [1.5, 2.5, 95]; cf[0, 3:] = np.eye(3).reshape(9)
    cm[1, :3] = [1, 2, 90]; cf[1, :3] = [3, 4, 99]                    # zero covariances: singular B
    prob = gpu_ctx.problem_create(cm, cf, [0, 1, 0], [0, 1, 0], [0, 3])
    pose = synth.pose_to_se2(0, 0, 0)
    r, J = prob.eval_emit(pose)
    assert np.isfinite(r[0]) and np.isnan(r[1]) and r[2] == r[0]
    assert gpu_ctx.take_bad_pairs() == 1
    out = capi.unpack_fused(prob.eval_fused(pose))
    assert np.isfinite(out["cost"][0]) and gpu_ctx.take_bad_pairs() == 1
    # identical distributions: r = 0, J defined as 0 (the reference's dual-number sqrt gives NaN here)
    prob2 = gpu_ctx.problem_create(cm[:1], cm[:1], [0], [0], [0, 1])
    r, J = prob2.eval_emit(pose)
    assert r[0] == 0.0 and np.all(J == 0.0)


def test_invalid_arguments(gpu_ctx):
    cm = np.zeros((1, 12), np.float32)
    with pytest.raises(capi.RandtError):
        gpu_ctx.problem_create(cm, cm, [0], [5], [0, 1])        # pair index out of range
    with pytest.raises(capi.RandtError):
        gpu_ctx.problem_create(cm, cm, [0], [0], [0, 2])        # seg_off does not span the pairs
    prob = gpu_ctx.problem_create(cm, cm, [0], [0], [0, 1])
    with pytest.raises(capi.RandtError):
        prob.eval_emit(np.zeros(3), variant=7)
    with pytest.raises(capi.RandtError):
        prob.eval_fused(np.array([1.0, 0, 0, 0]), capi.make_loss(capi.LOSS_BARRON, scale=-1.0))


def test_fused_writes_directly_into_pinned_host_memory(oracle, gpu_ctx):
    """pinned (mapped) result buffers are written by K3 itself (no staging copy): same bits as the staged path"""
    import ctypes
    case = H.make_registration_case(oracle, P.OXFORD, seed=9)
    cm, cf, im, jf = case["moving"]["cells"], case["fixed"]["cells"], case["im"], case["jf"]
    n = len(im)
    seg = [0, n // 4, n // 2, n]
    poses = np.stack([case["pose0"], synth.pose_to_se2(0.55, -0.35, 0.025), synth.pose_to_se2(0.45, -0.25, 0.015)])
    prob = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    loss = capi.make_loss(capi.LOSS_BARRON, 1.0, -2.0, 1.3, 0.5)
    staged = prob.eval_fused(poses, loss)                       # pageable numpy buffer -> device scratch + copy
    nbytes = 3 * capi.FUSED_STRIDE * 8
    ptr = capi.lib().randt_host_alloc(nbytes)
    assert ptr
    try:
        pinned = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_double)), shape=(3, capi.FUSED_STRIDE))
        pinned[:] = -1.0
        out = prob.eval_fused(poses, loss, out=pinned)
        assert out is pinned and np.array_equal(pinned, staged)
    finally:
        capi.lib().randt_host_free(ptr)


@pytest.mark.parametrize("with_empty", [False, True])
def test_fused_async_pipeline_matches_the_blocking_call(oracle, gpu_ctx, with_empty):
    """randt_eval_fused_async: six calls in flight over a two-slot input ring, each with its own poses / mu and result buffer; results
    equal the blocking call bit for bit (direct pinned stores without empty segments, staged copy with one), and pageable buffers
    are rejected"""
    case = H.make_registration_case(oracle, P.OXFORD, seed=12)
    cm, cf, im, jf = case["moving"]["cells"], case["fixed"]["cells"], case["im"], case["jf"]
    n = len(im)
    seg = [0, n // 5, n // 5, n // 2, n] if with_empty else [0, n // 5, n // 3, n // 2, n]
    S = len(seg) - 1
    prob = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    loss = capi.make_loss(capi.LOSS_BARRON, 1.0, -2.0, 1.0, 0.5)
    rng = np.random.default_rng(5)
    n_calls = 6
    # calls cycle through the three record layouts: full (24 doubles), packed (18: upper triangle of H), core (15: H, g, cost)
    stride = {0: capi.FUSED_STRIDE, 1: capi.PACKED_STRIDE, 2: capi.CORE_STRIDE}
    bufs = [(capi.PinnedArray((S, 4)), capi.PinnedArray((S,)), capi.PinnedArray((S, stride[i % 3]))) for i in range(n_calls)]
    try:
        for hp, hm, ho in bufs:
            hp.a[...] = np.stack([synth.pose_to_se2(*(np.array([0.5, -0.3, 0.02]) + rng.uniform(-0.05, 0.05, 3))) for _ in range(S)])
            hm.a[...] = rng.uniform(1.0, 4.0, S)
            ho.a[...] = -1.0
        for i, (hp, hm, ho) in enumerate(bufs):
            prob.eval_fused_async(hp.a, ho.a, loss, mu_per_seg=hm.a if i % 2 else None, packed=i % 3)
        # tickets: every call can be waited for on its own (calls complete in order)
        t_last = gpu_ctx.async_count()
        gpu_ctx.wait_async(t_last - 3)
        assert not np.any(bufs[n_calls - 4][2].a == -1.0)
        gpu_ctx.wait_async(t_last)
        assert not np.any(bufs[n_calls - 1][2].a == -1.0)
        with pytest.raises(capi.RandtError):
            gpu_ctx.wait_async(t_last + 1)
        gpu_ctx.sync()
        for i, (hp, hm, ho) in enumerate(bufs):
            want = prob.eval_fused(hp.a.copy(), loss, mu_per_seg=hm.a.copy() if i % 2 else None)
            assert np.array_equal(ho.a, want if i % 3 == 0 else capi.pack_fused(want)[:, :stride[i % 3]]), i
        with pytest.raises(capi.RandtError):
            prob.eval_fused_async(np.zeros((S, 4)), bufs[0][2].a, loss)
    finally:
        gpu_ctx.sync()
        for t in bufs:
            for b in t:
                b.close()


@pytest.mark.parametrize("k", [1, 2, 4])
def test_fused_ragged_segments_cover_every_chunk_shape(oracle, gpu_ctx, k):
    """segments of 1 ... 700 duos in one batch: single-chunk tiles, tiles ending exactly on a chunk, split chunks whose second tile also
    ends inside the chunk (several tiny segments in a row), multi-tile segments folded through partial records, empty segments"""
    rng = np.random.default_rng(100 + k)
    sizes = [1, 2, 31, 32, 33, 5, 63, 64, 65, 1, 1, 1, 40, 0, 255, 256, 257, 3, 600, 17, 96, 0, 700, 29, 30, 31, 32, 33, 34, 7]
    n_m = sum(sizes)
    cm = H.random_cells(rng, max(n_m, 1), extent=6.0); cf = H.random_cells(rng, 900, extent=6.0)
    im, jf, seg = [], [], [0]
    m = 0
    for sz in sizes:                      # sz moving cells, each with k neighbours (k consecutive pairs share their moving cell)
        for _ in range(sz):
            for j in rng.choice(900, k, replace=False):
                im.append(m); jf.append(int(j))
            m += 1
        seg.append(len(im))
    im = np.array(im, np.uint32); jf = np.array(jf, np.uint32)
    S = len(sizes)
    poses = np.stack([synth.pose_to_se2(*rng.uniform(-0.3, 0.3, 3)) for _ in range(S)])
    mus = rng.uniform(1.0, 3.0, S)
    loss = capi.make_loss(capi.LOSS_BARRON, 1.5, -2.0, 1.0, 0.02)
    prob = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    out = capi.unpack_fused(prob.eval_fused(poses, loss, mu_per_seg=mus))
    fo, _ = oracle.fused_batch(0, cm, cf, im, jf, seg, poses, (loss.kind, loss.scale, loss.alpha, 1.0, loss.weight), mu_per_seg=mus)
    for s_ in range(S):
        if sizes[s_] == 0:
            assert out["n"][s_] == 0 and out["cost"][s_] == 0 and np.all(out["H"][s_] == 0)
            continue
        assert out["n"][s_] == sizes[s_] * k
        assert H.rel_err(out["H"][s_], fo["H"][s_]) < 1e-9, (s_, sizes[s_])
        assert H.rel_err(out["g"][s_], fo["g"][s_]) < 1e-9, (s_, sizes[s_])
        assert abs(out["cost"][s_] - fo["cost"][s_]) <= 1e-10 * abs(fo["cost"][s_])
        assert abs(out["max_r"][s_] - fo["max_r"][s_]) <= 1e-12 * fo["max_r"][s_]
    # EMIT walks the unsplit schedule over the same record table
    r, J = prob.eval_emit(poses)
    for s_ in (0, 4, 18, 22):
        a, b = seg[s_], seg[s_ + 1]
        ro, Jo = oracle.eval_pairs(0, cm, cf, im[a:b], jf[a:b], poses[s_], 0)
        assert np.max(np.abs(r[a:b] - ro) / ro) < TOL and rowwise(J[a:b], Jo) < 1e-8


def test_a_warp_walks_its_descriptor_queue_through_several_refills(oracle, gpu_ctx, monkeypatch):
    """ADVICE (round 1): a warp that owns more than 64 chunks refills its 2 x 32-entry descriptor queue while chunks are in flight.  With
    the schedule built for 4 warps (RANDT_K3_MAX_WARPS) every warp owns ~170 chunks (five refills): FUSED and EMIT must equal the
    default schedule's results bit for bit (a segment's sums depend on its own tiles only) and the oracle to tolerance."""
    rng = np.random.default_rng(77)
    sizes = [500] * 20 + [37, 1, 700, 64, 500, 500, 33, 500] + [500] * 14
    n_m = sum(sizes)
    cm = H.random_cells(rng, n_m, extent=6.0); cf = H.random_cells(rng, 900, extent=6.0)
    im = np.repeat(np.arange(n_m, dtype=np.uint32), 2); jf = rng.integers(0, 900, 2 * n_m).astype(np.uint32)
    seg = np.concatenate([[0], 2 * np.cumsum(sizes)]).astype(np.uint32)
    S = len(sizes)
    poses = np.stack([synth.pose_to_se2(*rng.uniform(-0.3, 0.3, 3)) for _ in range(S)])
    loss = capi.make_loss(capi.LOSS_BARRON, 1.5, -2.0, 1.4, 0.02)
    ref = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    monkeypatch.setenv("RANDT_K3_MAX_WARPS", "4")
    few = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    monkeypatch.delenv("RANDT_K3_MAX_WARPS")
    sch = few.schedule()
    assert sch["n_warps"] == 4 and ref.schedule()["n_warps"] > 4
    assert np.min(np.diff(sch["woff_a"])) > 128 and np.min(np.diff(sch["woff_b"])) > 128        # > 2 queues' worth of chunks per warp
    a = few.eval_fused(poses, loss); b = ref.eval_fused(poses, loss)
    assert np.array_equal(a, b)
    ra, Ja = few.eval_emit(poses); rb, Jb = ref.eval_emit(poses)
    assert np.array_equal(ra, rb) and np.array_equal(Ja, Jb)
    out = capi.unpack_fused(a)
    fo, _ = oracle.fused_batch(0, cm, cf, im, jf, seg, poses, (loss.kind, loss.scale, loss.alpha, loss.mu, loss.weight))
    for s_ in range(S):
        assert H.rel_err(out["H"][s_], fo["H"][s_]) < 1e-9 and H.rel_err(out["g"][s_], fo["g"][s_]) < 1e-9
    few.close(); ref.close()


def test_basis_record_is_the_normal_equations_before_the_chain_rule(oracle, gpu_ctx):
    """packed == 3 (RANDT_BASIS_*, 80 B per pose): H_b, g_b, cost in the functor's (theta, tx, ty) basis; expanded with F^T . F it equals
    the ambient core record (packed == 2) of the same evaluation, projected with Sophus' PlusJacobian it equals the tangent-space system
    the oracle builds from its ambient sums; ragged segments incl. multi-tile ones (partial fold) and poses off the unit circle"""
    rng = np.random.default_rng(31)
    sizes = [40, 1, 300, 33, 700, 64, 5]
    n_m = sum(sizes)
    cm = H.random_cells(rng, n_m, extent=6.0); cf = H.random_cells(rng, 500, extent=6.0)
    im = np.repeat(np.arange(n_m, dtype=np.uint32), 2); jf = rng.integers(0, 500, 2 * n_m).astype(np.uint32)
    seg = np.concatenate([[0], 2 * np.cumsum(sizes)]).astype(np.uint32)
    S = len(sizes)
    poses = np.stack([synth.pose_to_se2(*rng.uniform(-0.3, 0.3, 3)) for _ in range(S)])
    poses[2, :2] *= 1.003; poses[5, :2] *= 0.998
    loss = capi.make_loss(capi.LOSS_BARRON, 1.5, -2.0, 1.4, 0.02)
    prob = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    hp = capi.PinnedArray((S, 4)); hp.a[...] = poses
    hb = capi.PinnedArray((S, capi.BASIS_STRIDE)); hc = capi.PinnedArray((S, capi.CORE_STRIDE))
    prob.eval_fused_async(hp.a, hc.a, loss, packed=2); gpu_ctx.sync()
    prob.eval_fused_async(hp.a, hb.a, loss, packed=3); gpu_ctx.sync()
    core = capi.basis_to_core(hb.a, poses)
    assert np.array_equal(core[:, 14], hc.a[:, 14])                                 # the cost slot travels unchanged
    scale = np.max(np.abs(hc.a[:, :14]), axis=1, keepdims=True)
    assert np.max(np.abs(core[:, :14] - hc.a[:, :14]) / scale) < 1e-14
    # the manifold system: Q^T H_b Q with Q = [[0, 0, 1], [c, -s, 0], [s, c, 0]] for unit (c, s) == PlusJacobian^T H_ambient PlusJacobian
    fo, _ = oracle.fused_batch(0, cm, cf, im, jf, seg, poses, (loss.kind, loss.scale, loss.alpha, loss.mu, loss.weight))
    for s_ in (0, 1, 3, 4, 6):
        c, s = poses[s_, 0], poses[s_, 1]
        Q = np.array([[0, 0, 1.0], [c, -s, 0], [s, c, 0]]); Pj = np.array([[0, 0, -s], [0, 0, c], [c, -s, 0], [s, c, 0]])
        Hb = np.zeros((3, 3)); iu = np.triu_indices(3); Hb[iu] = hb.a[s_, :6]; Hb = Hb + Hb.T - np.diag(np.diag(Hb))
        assert H.rel_err(Q.T @ Hb @ Q, Pj.T @ fo["H"][s_] @ Pj) < 1e-9 and H.rel_err(Q.T @ hb.a[s_, 6:9], Pj.T @ fo["g"][s_]) < 1e-9
    # not defined for the four-dimensional basis of the 2-D SE(2) functor, nor without a Jacobian evaluation
    with pytest.raises(capi.RandtError):
        prob.eval_fused_async(hp.a, hb.a, loss, packed=3, variant=capi.VAR_SE2_XY)
    with pytest.raises(capi.RandtError):
        prob.eval_fused_async(hp.a, hb.a, loss, packed=3, want_jac=False)
    gpu_ctx.sync()
    for b in (hp, hb, hc):
        b.close()
    prob.close()


def test_fused_one_very_long_segment(oracle, gpu_ctx):
    """a single segment of 300 k pairs: ~590 tiles folded by the ticket/partial path"""
    rng = np.random.default_rng(5)
    cm = H.random_cells(rng, 4000, extent=10.0); cf = H.random_cells(rng, 6000, extent=10.0)
    P_ = 300000
    im = np.sort(rng.integers(0, 4000, P_)).astype(np.uint32); jf = rng.integers(0, 6000, P_).astype(np.uint32)
    pose = synth.pose_to_se2(0.1, -0.2, 0.03)
    loss = capi.make_loss(capi.LOSS_BARRON, 2.0, -1.0, 1.7, 1e-3)
    prob = gpu_ctx.problem_create(cm, cf, im, jf, [0, P_])
    o = capi.unpack_fused(prob.eval_fused(pose, loss))
    fo = oracle.fused(0, cm, cf, im, jf, pose, (loss.kind, loss.scale, loss.alpha, loss.mu, loss.weight), True)
    assert o["n"][0] == P_
    assert H.rel_err(o["H"][0], fo["H"]) < 1e-9 and H.rel_err(o["g"][0], fo["g"]) < 1e-9
    assert abs(o["cost"][0] - fo["cost"]) <= 1e-10 * fo["cost"]


def test_compact_records_hold_asymmetric_cells_and_overflow(oracle, gpu_ctx):
    """K3 streams 112-byte records that carry, per off-diagonal couple (a, b) of a covariance, float_rz(a + b) and a 2-bit code (common.cuh).
    Covariances that are asymmetric at float-ulp level (what V L V^-1 and R S R^T leave behind) must be held exactly; a couple whose two
    entries are far apart in exponent (rounding noise around zero) cannot be and must take the full-precision overflow record."""
    rng = np.random.default_rng(77)
    cm = H.random_cells(rng, 300); cf = H.random_cells(rng, 700)
    for c in (cm, cf):
        for i, j in ((4, 6), (5, 9), (8, 10)):
            steps = rng.integers(-3, 4, len(c))
            v = c[:, j].copy()
            for _ in range(3):
                up = steps > 0; dn = steps < 0
                v[up] = np.nextafter(v[up], np.float32(np.inf)); v[dn] = np.nextafter(v[dn], np.float32(-np.inf))
                steps -= np.sign(steps)
            c[:, j] = v
    esc_m = np.array([5, 77, 201]); esc_f = np.array([0, 13, 400, 699])
    cm[esc_m, 4] = 1.0e-12; cm[esc_m, 6] = 3.1e-17
    cf[esc_f, 5] = -2.5e-11; cf[esc_f, 9] = 7.7e-16
    im = np.repeat(np.arange(300, dtype=np.uint32), 3)                    # three neighbours per moving cell: duos of 2 + 1
    jf = rng.integers(0, 700, len(im)).astype(np.uint32)
    jf[:8] = [0, 13, 400, 699, 0, 13, 400, 699]
    seg = [0, 450, len(im)]
    poses = np.stack([synth.pose_to_se2(0.3, -0.2, 0.05), synth.pose_to_se2(-0.4, 0.1, -0.3)])
    prob = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    n_duos, rec_bytes, n_ovf = prob.layout()
    assert rec_bytes == 112 and n_duos == 600
    touched = np.isin(im, esc_m) | np.isin(jf, esc_f)
    want_ovf = len(np.unique((np.arange(len(im)) // 3 * 2 + (np.arange(len(im)) % 3) // 2)[touched]))
    assert n_ovf == want_ovf and 0 < n_ovf < n_duos // 4
    r, J = prob.eval_emit(poses)
    loss = LOSSES[1]
    out = capi.unpack_fused(prob.eval_fused(poses, loss))
    for s in range(2):
        a, b = seg[s], seg[s + 1]
        ro, Jo = oracle.eval_pairs(0, cm, cf, im[a:b], jf[a:b], poses[s], 0)
        assert np.max(np.abs(r[a:b] - ro) / ro) < TOL
        assert rowwise(J[a:b], Jo) < 1e-8
        fo = oracle.fused(0, cm, cf, im[a:b], jf[a:b], poses[s], loss_tuple(loss), True)
        assert abs(out["cost"][s] - fo["cost"]) < TOL * abs(fo["cost"])
        assert H.rel_err(out["H"][s], fo["H"]) < 1e-8 and H.rel_err(out["g"][s], fo["g"]) < 1e-8
    assert gpu_ctx.take_bad_pairs() == 0


@pytest.mark.parametrize("shape", ["ragged", "bench-like", "few long", "one"])
def test_device_built_schedule_equals_the_host_builder(oracle, gpu_ctx, shape):
    """The host computes the tile assignment only; the chunk lists of both plans are written by a kernel running the same walk
    (schedule.hpp walk_warp).  They must be the lists the all-host builder (which tests/test_schedule_cpu.py checks on the CPU) makes."""
    from randt_slam_b200 import hostapi
    rng = np.random.default_rng(5)
    if shape == "ragged":
        sizes = [1, 2, 31, 32, 33, 5, 63, 64, 65, 1, 1, 1, 40, 0, 255, 256, 257, 3, 600, 17, 96, 0, 700, 29, 30, 31, 32, 33, 34, 7]
    elif shape == "bench-like":
        sizes = list(rng.integers(60, 130, 3000))
    elif shape == "few long":
        sizes = list(rng.integers(2000, 9000, 12))
    else:
        sizes = [180]
    k = 2
    n_m = int(sum(sizes))
    cm = H.random_cells(rng, max(n_m, 1), extent=6.0); cf = H.random_cells(rng, 64, extent=6.0)
    im = np.repeat(np.arange(n_m, dtype=np.uint32), k)          # every moving cell once, with k neighbours: one duo each
    jf = rng.integers(0, 64, n_m * k).astype(np.uint32)
    seg = np.concatenate([[0], np.cumsum(np.asarray(sizes, np.int64) * k)]).astype(np.uint32)
    prob = gpu_ctx.problem_create(cm, cf, im, jf, seg)
    dev = prob.schedule()
    assert list(np.diff(dev["duo_off"])) == list(sizes)
    host = hostapi.build_schedule(dev["duo_off"], dev["warp_budget"])
    assert dev["n_warps"] == host["n_warps"] and dev["n_tiles"] == len(host["tiles"])
    assert np.array_equal(dev["woff_a"], host["woff_a"]) and np.array_equal(dev["woff_b"], host["woff_b"])
    assert np.array_equal(dev["plan_a"], host["plan_a"])
    assert np.array_equal(dev["plan_b"], host["plan_b"])
    prob.close()
