"""Device sparse pose adjustment (csrc/dl_posegraph.cu; reference OptimizationProblem3D::Solve, optimization_problem_3d.cc:259-589)
against the oracle (oracle/orc_posegraph.h, pinned to the reference's ReducesNoise test): same problems, same LM trajectory
(iteration counts), poses to 1e-8. With a communicator the normal equations go through ncclAllReduce(fp64) (one rank here; N
ranks in tools/bench_pose_graph.py)."""
import numpy as np
import pytest

from test_posegraph_oracle import aa_to_q, angle, compose, inverse, qmul

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import dliom
    c = dliom.Context(0)
    yield c
    c.close()


def close(a, b, tol_t=1e-8, tol_r=1e-8):
    for p, q in zip(a, b):
        rel = compose(inverse(np.asarray(q)), np.asarray(p))
        assert np.linalg.norm(rel[:3]) < tol_t and angle(rel) < tol_r, (p, q)


def test_exact_recovery_matches_oracle(ctx, orc):
    rng = np.random.default_rng(2)
    submaps = [np.array([0, 0, 0, 1.0, 0, 0, 0]), np.array([4.0, 1.0, 0.2, *aa_to_q([0, 0, 0.5])])]
    truth = [np.array([*rng.uniform(-8, 8, 3), *aa_to_q(rng.uniform(-0.6, 0.6, 3))]) for _ in range(12)]
    cons = [(s, n, compose(inverse(submaps[s]), truth[n]), 1.0, 1.0) for s in range(2) for n in range(12)]
    start_nodes = [compose(t, np.array([*rng.uniform(-0.5, 0.5, 3), *aa_to_q(rng.uniform(-0.2, 0.2, 3))])) for t in truth]
    start_submaps = [submaps[0], compose(submaps[1], np.array([0.3, -0.2, 0.1, *aa_to_q([0.02, -0.03, 0.1])]))]
    ws, wn, wsum = orc.pose_graph_solve(start_submaps, start_nodes, cons)
    gs, gn, gsum, info = ctx.pose_graph_solve(start_submaps, start_nodes, cons)
    assert info.num_local_parameters == 2 + 6 * 13 and info.all_reduce_count == 0
    assert gsum["num_iterations"] == wsum["num_iterations"] and gsum["termination"] == wsum["termination"] == 0
    assert gsum["final_cost"] < 1e-12 * max(gsum["initial_cost"], 1.0)
    assert abs(gsum["initial_cost"] - wsum["initial_cost"]) <= 1e-12 * wsum["initial_cost"]
    close(gs, ws)
    close(gn, wn)
    assert np.array_equal(gs[0][:3], start_submaps[0][:3])     # the first submap's translation is a constant block


def test_reduces_noise_like_the_reference_and_the_oracle(ctx, orc):
    """optimization_problem_3d_test.cc:106-196 with the oracle's construction (tests/test_posegraph_oracle.py)."""
    rng = np.random.default_rng(0)

    def random_transform(ts, rs):
        return np.array([*rng.uniform(-ts, ts, 3), *aa_to_q(rng.uniform(-rs, rs, 3))])

    def random_yaw_only(ts, rs):
        return np.array([*rng.uniform(-ts, ts, 3), *aa_to_q([0, 0, rng.uniform(-rs, rs)])])

    def add_noise(t, noise):
        return np.array([*(t[:3] + noise[:3]), *qmul(noise[3:], t[3:])])
    n = 100
    truth = [random_transform(10.0, 3.0) for _ in range(n)]
    noise = [random_yaw_only(0.2, 0.3) for _ in range(n)]
    nodes = [add_noise(t, z) for t, z in zip(truth, noise)]
    submap2 = np.array([0, 0, 0, *aa_to_q([0, 0, np.pi])])
    cons = []
    for j in range(n):
        cons.append((0, j, add_noise(truth[j], noise[j]), 1.0, 1.0))
        cons.append((1, j, add_noise(truth[j], random_yaw_only(0.2, 0.3)), 1.0, 1.0))
        cons.append((2, j, compose(compose(inverse(submap2), truth[j]), random_transform(1e3, 3.0)), 1e-9, 1e-9))
    ident = np.array([0, 0, 0, 1.0, 0, 0, 0])

    def errors(ps):
        return (sum(np.linalg.norm(t[:3] - p[:3]) for t, p in zip(truth, ps)),
                sum(angle(compose(inverse(t), p)) for t, p in zip(truth, ps)))
    t_before, r_before = errors(nodes)
    ws, wn, wsum = orc.pose_graph_solve([ident, ident, submap2], nodes, cons, max_iter=50)
    gs, gn, gsum, info = ctx.pose_graph_solve([ident, ident, submap2], nodes, cons, max_iter=50)
    t_after, r_after = errors(gn)
    assert 0.8 * t_before > t_after and 0.8 * r_before > r_after          # the reference's assertion
    assert gsum["num_iterations"] == wsum["num_iterations"]
    assert abs(gsum["final_cost"] - wsum["final_cost"]) <= 1e-9 * wsum["final_cost"]
    close(gn, wn, 1e-6, 1e-6)
    assert info.num_local_parameters == 2 + 6 * 102


def test_fix_z_and_argument_checks(ctx, orc):
    import dliom
    rng = np.random.default_rng(4)
    ident = np.array([0, 0, 0, 1.0, 0, 0, 0])
    truth = [np.array([*rng.uniform(-5, 5, 3), *aa_to_q(rng.uniform(-0.3, 0.3, 3))]) for _ in range(6)]
    lifted = [t + np.array([0.2, -0.1, 0.4, 0, 0, 0, 0]) for t in truth]
    cons = [(0, k, truth[k], 1.0, 1.0) for k in range(6)]
    _, wn, wsum = orc.pose_graph_solve([ident], lifted, cons, fix_z=True)
    _, gn, gsum, _ = ctx.pose_graph_solve([ident], lifted, cons, fix_z=True)
    assert all(a[2] == b[2] for a, b in zip(gn, lifted))
    close(gn, wn, 1e-7, 1e-7)
    assert gsum["num_iterations"] == wsum["num_iterations"]
    with pytest.raises(dliom.DlError):
        ctx.pose_graph_solve([ident], lifted, [(0, 9, truth[0], 1.0, 1.0)])      # node outside the graph
    with pytest.raises(dliom.DlError):
        ctx.pose_graph_solve([ident], [ident] * 600, [])                          # beyond the dense solver's size


def test_sharded_constraints_through_the_all_reduce(ctx, orc):
    """Constraints split over the ranks: with one rank the all-reduce is the identity, so the result must equal the comm-less solve
    bit for bit, and the communication bookkeeping must show one reduction of (n^2 + n + 1) doubles per evaluation."""
    import dliom
    comm = dliom.Comm(ctx, dliom.comm_unique_id(), 0, 1)
    rng = np.random.default_rng(8)
    ident = np.array([0, 0, 0, 1.0, 0, 0, 0])
    submaps = [ident, np.array([3.0, -1.0, 0.1, *aa_to_q([0, 0, -0.4])])]
    truth = [np.array([*rng.uniform(-6, 6, 3), *aa_to_q(rng.uniform(-0.5, 0.5, 3))]) for _ in range(20)]
    cons = [(s, n, compose(compose(inverse(submaps[s]), truth[n]), np.array([*rng.normal(0, 0.02, 3), *aa_to_q(rng.normal(0, 0.01, 3))])),
             1.1e4 ** 0.5, 1e5 ** 0.5) for s in range(2) for n in range(20)]
    start = [compose(t, np.array([*rng.uniform(-0.3, 0.3, 3), *aa_to_q(rng.uniform(-0.1, 0.1, 3))])) for t in truth]
    a = ctx.pose_graph_solve(submaps, start, cons)
    b = ctx.pose_graph_solve(submaps, start, cons, comm=comm)
    assert a[2]["num_iterations"] == b[2]["num_iterations"]
    close(a[1], b[1], 1e-9, 1e-9)        # fp64 atomics order the local sums differently from run to run: last-bit differences
    n_local = 2 + 6 * 21
    assert b[3].all_reduce_count == b[2]["num_evaluations"] and b[3].all_reduce_bytes == 8 * (n_local * n_local + n_local + 1)
    assert b[3].all_reduce_ms > 0
    w = orc.pose_graph_solve(submaps, start, cons)
    close(b[1], w[1], 1e-7, 1e-7)
    comm.close()
