"""Host-side problem tables (quadruped_landing_b200/problem.py) against the oracle and the notebook."""
import numpy as np

from oracle import oracle
import quadruped_landing_b200 as ql


def test_lqr_cost_matches_oracle_bitwise():
    rng = np.random.default_rng(3)
    for _ in range(50):
        Q, R = rng.uniform(0, 20, 15), rng.uniform(0, 1, 5)
        xf, uf = rng.normal(size=15), rng.normal(size=5) * 50
        c = ql.LQRCost(np.diag(Q), np.diag(R), xf, uf)
        q0, r0, c0 = oracle.lqr_cost(Q, R, xf, uf)
        assert np.array_equal(c.q, q0) and np.array_equal(c.r, r0) and c.c == c0


def test_default_instance_values():
    p = ql.default_problem()
    # main.ipynb:92-93: v_init_y = sqrt(2*9.81*2) printed as 6.26418390534633
    assert p.x0[8] == -6.26418390534633 and p.x0[2] == -30 * np.pi / 180 and p.x0[9] == -np.pi / 2
    assert p.xf[1] == np.sqrt(0.25 ** 2 + 0.25 ** 2) and p.xf[0] == -0.25 and p.xf[5] == -0.5
    assert (p.N, p.k_trans, p.init_mode) == (61, 21, 1)
    # only two distinct stage-cost structs exist because Q[15] = 0
    assert np.array_equal(p.r[0], p.r[19]) and np.array_equal(p.r[20], p.r[59]) and not np.array_equal(p.r[0], p.r[20])
    assert p.R[60].max() == 0.0 and p.c[60] == ql.LQRCost(p.Q[60], p.R[60], p.xf).c


def test_reference_trajectory_layout():
    m = ql.PlanarQuadruped()
    _, xt = ql.default_states(m)
    for mode in (1, 2):
        X, U = ql.reference_trajectory(m, 61, 21, xt, mode, 0.009)
        assert len(X) == 61 and len(U) == 60
        lead, other = (1, 3) if mode == 1 else (3, 1)
        assert U[0][lead] == 98.10000000000001 and U[0][other] == 0.0 and U[0][4] == 0.001       # ref_traj.jl:24,35
        assert U[20][1] == U[20][3] == 49.050000000000004 and U[20][4] == 0.02                     # ref_traj.jl:26-27,36
        assert abs(X[60][14] - 0.54) < 1e-15 and np.array_equal(X[5][:14], xt[:14])


def test_initial_guess_layout():
    p = ql.default_problem()
    z = ql.initial_guess(p)
    X, U = ql.unpackZ(61, z)
    assert np.array_equal(X[0][:14], p.x0[:14]) and np.allclose(X[20][:14], p.xf[:14], atol=1e-15)
    assert abs(X[20][14] - 0.02) < 1e-15 and abs(X[60][14] - 0.82) < 1e-13        # 20*0.001 + 40*0.02
    assert U[0][1] == 98.10000000000001


def test_sweep_and_batched_guess_match_the_single_problem_formulas():
    p = ql.default_problem()
    x0 = ql.sweep_initial_states(p.model, [0.25, 2.0, 3.0], [-40.0, -30.0, -5.0])
    assert x0.shape == (9, 15) and np.array_equal(x0[4], p.x0)                       # h=2, theta=-30 is the notebook case
    assert x0[0][8] == -np.sqrt(2 * 9.81 * 0.25) and x0[2][2] == -5.0 * np.pi / 180  # main.ipynb:92-93,118,122
    Z = ql.initial_guess_batch(p, x0)
    for b in range(9):
        assert np.array_equal(Z[b], ql.initial_guess(ql.build_problem(xinit=x0[b])))


def test_solution_csv_round_trip(tmp_path, golden):
    path = str(tmp_path / "z.csv")
    ql.save_solution_csv(path, golden["data_6"])
    assert np.array_equal(ql.load_solution_csv(path), golden["data_6"])              # writedlm / loadtxt format
    tab = ql.solution_table(golden["data_6"], 61)
    assert tab.shape == (61, 20) and tab[60, 14] == 0.848539898959304 and not tab[60, 15:].any()
    assert tab[59, 16] + tab[59, 18] == 98.10000000000001                            # plot_data.py columns 15-18 = forces


def test_ragged_offsets_are_aligned_and_pack_round_trips():
    """Host logic of the ragged path (no GPU needed): rows are padded to an even length so every row of the flat
    arrays starts 16-byte aligned; pack() places the vectors at z_off."""
    probs = [ql.build_problem(N=N, k_trans=kt, init_mode=im) for N, kt, im in [(31, 11, 1), (41, 14, 2), (61, 21, 1)]]
    ev = ql.RaggedEvaluator(probs)
    rng = np.random.default_rng(0)
    class_of = rng.integers(0, 3, size=50)
    off = ev.offsets(class_of)
    for key, widths in (("z_off", ev.n), ("g_off", ev.m), ("j_off", ev.nnz)):
        o = off[key]
        assert o[0] == 0 and np.all(o % 2 == 0)
        assert np.all(np.diff(o) >= widths[class_of]) and np.all(np.diff(o) - widths[class_of] <= 1)
    vecs = [rng.normal(size=probs[c].n_nlp) for c in class_of]
    Z = ev.pack(class_of, vecs)
    assert Z.shape == (off["z_off"][-1],)
    for b, v in enumerate(vecs):
        assert np.array_equal(Z[off["z_off"][b]:off["z_off"][b] + len(v)], v)
