"""The opt-in kinematic (leg-length) rows: commented out in the reference (nlp.jl:60,70; constraints.jl:115-138,
276-288), switched on by QLNLP_WITH_KINEMATICS.  Default off must leave everything as it was; on, the structure is
the oracle's (mask-derived) structure and the GPU values equal the oracle's bit for bit."""
import numpy as np
import pytest

from conftest import perturbed_batch
from oracle.oracle import Oracle
import quadruped_landing_b200 as ql

CLASSES = [(61, 21, 1), (31, 11, 2), (2, 1, 1), (33, 33, 2), (65, 2, 1)]


@pytest.mark.parametrize("N,kt,im", CLASSES)
@pytest.mark.parametrize("pattern", ["block", "true"])
def test_dimensions_structure_and_bounds(N, kt, im, pattern):
    p = ql.build_problem(N=N, k_trans=kt, init_mode=im)
    off = ql.HybridNLP.from_problem(p, pattern=pattern)
    on = ql.HybridNLP.from_problem(p, pattern=pattern, kinematics=True)
    o = Oracle(p, kinematics=True)
    assert off.m_nlp == 18 * N - kt + 16 and on.m_nlp == off.m_nlp + 2 * N == o.m_nlp
    assert on.nnz == off.nnz + 8 * N and on.nnz_block == off.nnz_block + 8 * N == o.nnz
    assert len(on.cinds) == 8 and on.cinds[7][0] == off.m_nlp + 1 and on.cinds[7][-1] == on.m_nlp
    r, c = on.jacobian_structure_arrays()
    r0, c0 = o.jacobian_structure() if pattern == "block" else o.jacobian_structure_true()
    assert np.array_equal(r, r0) and np.array_equal(c, c0)
    lb, ub = on.constraint_bounds()
    lb0, ub0 = o.constraint_bounds()
    assert np.array_equal(lb, lb0) and np.array_equal(ub, ub0)
    assert ub[-1] == p.model.l1 + p.model.l2 + p.model.lb / 2 and lb[-1] == 0.0
    # the dense grid grows with m_nlp
    d = ql.HybridNLP.from_problem(p, use_sparse_jacobian=False, kinematics=True)
    assert d.nnz == on.m_nlp * on.n_nlp


@pytest.mark.gpu
@pytest.mark.parametrize("N,kt,im", CLASSES)
def test_gpu_values_equal_the_oracle(N, kt, im):
    torch = pytest.importorskip("torch")
    p = ql.build_problem(N=N, k_trans=kt, init_mode=im)
    o = Oracle(p, kinematics=True)
    base = ql.initial_guess(p) if kt > 1 else np.zeros(p.n_nlp)
    B = 130
    Z = perturbed_batch(p, [base], B, 5e-2, 17)
    for pattern in ("block", "true"):
        nlp = ql.HybridNLP.from_problem(p, pattern=pattern, kinematics=True)
        ref = o.eval_batch(Z, pattern=pattern)
        out = nlp.eval_batch(torch.from_numpy(Z).cuda())
        torch.cuda.synchronize()
        kin_g = slice(nlp.m_nlp - 2 * N, nlp.m_nlp)
        g = out["g"].cpu().numpy()
        assert np.array_equal(g[:, kin_g], ref["g"][:, kin_g])                      # sqrt / IEEE division: same bits
        rows, cols = nlp.jacobian_structure_arrays()
        kin_j = rows > nlp.m_nlp - 2 * N
        jac = out["jac"].cpu().numpy()
        assert kin_j.sum() == 8 * N and np.array_equal(jac[:, kin_j], ref["jac"][:, kin_j])
        # everything else: the same values the handle without the flag produces
        plain = ql.HybridNLP.from_problem(p, pattern=pattern).eval_batch(torch.from_numpy(Z).cuda())
        torch.cuda.synchronize()
        assert np.array_equal(jac[:, ~kin_j], plain["jac"].cpu().numpy())
        assert np.array_equal(g[:, :nlp.m_nlp - 2 * N], plain["g"].cpu().numpy())
        assert torch.equal(out["f"], plain["f"]) and torch.equal(out["grad"], plain["grad"])
        # host-pointer batches and the MOI callbacks take the same route
        host = nlp.eval_batch_host(Z[:70])
        for k in ("f", "grad", "g", "jac"):
            assert np.array_equal(host[k], out[k][:70].cpu().numpy()), k
        vals, gg = np.empty(nlp.nnz), np.empty(nlp.m_nlp)
        nlp.eval_constraint_jacobian(vals, Z[3])
        nlp.eval_constraint(gg, Z[3])
        assert np.array_equal(vals, jac[3]) and np.array_equal(gg, g[3])
        assert nlp.eval_objective(Z[3]) == float(out["f"][3])
    # the reference's dense structure with the extra rows
    dn = ql.HybridNLP.from_problem(p, use_sparse_jacobian=False, kinematics=True)
    vec = np.full(dn.nnz, np.nan)
    dn.eval_constraint_jacobian(vec, Z[0])
    dense = vec.reshape(dn.n_nlp, dn.m_nlp).T
    want = np.zeros_like(dense)
    rb, cb = o.jacobian_structure()
    want[rb - 1, cb - 1] = o.eval_batch(Z[:1])["jac"][0]
    mask_trig = np.zeros_like(dense, dtype=bool)
    mask_trig[dn.m_nlp - 3 * N:dn.m_nlp - 2 * N, 2::20] = True        # body-clearance d/dtheta: cos of CUDA's libm
    assert np.array_equal(dense[~mask_trig], want[~mask_trig])
    assert np.abs(dense[mask_trig] - want[mask_trig]).max() <= 4 * np.finfo(np.float64).eps
    # what the flag does not cover fails loudly
    kh = ql.HybridNLP.from_problem(p, kinematics=True, hessian=True)
    with pytest.raises(ql.QlnlpError):
        kh.eval_hessian_lagrangian(np.empty(kh.nnz_hess), Z[0], 1.0, np.zeros(kh.m_nlp))
