"""CPU tests: the numpy oracle against (1) the reference's own golden fixtures,
(2) outputs of the live reference (committed), (3) the live reference itself
when /root/reference is present (build container only)."""
import numpy as np
import pytest

from oracle import mca_oracle as orc
from oracle.ref_harness import reference_available

RTOL = 1e-3      # the reference's own test tolerance (test_integration_xarray.py:33-35)
ATOL = 1e-3


def _solve_fixture(fixtures, cplx):
    m = orc.make_model(fixtures["sst"].copy(), fixtures["prcp"].copy())
    return orc.solve(m, complexify=cplx)


@pytest.mark.parametrize("cplx", [False, True])
def test_oracle_matches_reference_fixture_sigma(fixtures, cplx):
    m = _solve_fixture(fixtures, cplx)
    gold = fixtures["sv_cplx" if cplx else "sv_std"]
    assert m.rank == 155
    np.testing.assert_allclose(m.sigma[:100], gold[:100], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("cplx,nm", [(False, 100), (True, 40)])
def test_oracle_matches_reference_fixture_eofs(fixtures, cplx, nm):
    m = _solve_fixture(fixtures, cplx)
    tag = "cplx" if cplx else "std"
    got = orc.eofs(m, nm, rotated=False)
    gl, gr = fixtures["eofs_%s_sst" % tag], fixtures["eofs_%s_prcp" % tag]
    al, ar = orc.align_modes(gl, got["left"], got["right"])
    assert np.array_equal(np.isnan(al), np.isnan(gl))
    # the trailing well-separated modes; near-degenerate pairs may mix
    np.testing.assert_allclose(np.nan_to_num(al)[..., :20], np.nan_to_num(gl)[..., :20],
                               rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(np.nan_to_num(ar)[..., :20], np.nan_to_num(gr)[..., :20],
                               rtol=RTOL, atol=ATOL)


def test_analytic_signal_matches_scipy():
    from scipy.signal import hilbert
    rng = np.random.default_rng(0)
    for n in (16, 17, 492):
        x = rng.standard_normal((n, 7)).astype(np.float32)
        np.testing.assert_allclose(orc.analytic_signal(x), hilbert(x, axis=0), rtol=1e-5, atol=1e-5)
        x64 = x.astype(np.float64)
        np.testing.assert_allclose(orc.analytic_signal(x64), hilbert(x64, axis=0), rtol=1e-12, atol=1e-12)


def _check_rotated(live, tag, m, n):
    np.testing.assert_allclose(orc.get_variance(m, n), live[tag + "/variance"], rtol=1e-6)
    np.testing.assert_array_equal(m.var_idx, live[tag + "/var_idx"])
    got_e, got_p = orc.eofs(m, n), orc.pcs(m, n)
    keys = m.keys
    ge = [got_e[k] for k in keys] + [got_p[k] for k in keys]
    re = [live[tag + "/eofs_" + k] for k in keys] + [live[tag + "/pcs_" + k] for k in keys]
    al = orc.align_modes(re[0], ge[0], *ge[1:])
    for a, r in zip(al, re):
        np.testing.assert_allclose(np.nan_to_num(a), np.nan_to_num(r), rtol=1e-4, atol=1e-5)


def test_oracle_matches_live_case_A(live):
    m = orc.solve(orc.make_model(live["A/left"].copy(), live["A/right"].copy()))
    np.testing.assert_allclose(m.sigma, live["A/sigma"], rtol=2e-5, atol=1e-5)
    np.testing.assert_allclose([m.total_covariance, m.total_squared_covariance], live["A/total"], rtol=1e-5)
    orc.rotate(m, 8, 1)
    _check_rotated(live, "A/varimax", m, 8)
    m = orc.solve(orc.make_model(live["A/left"].copy(), live["A/right"].copy()))
    orc.rotate(m, 8, 2)
    _check_rotated(live, "A/promax2", m, 8)
    np.testing.assert_allclose(orc.rotation_matrix(m, True) @ m.R.conj().T, np.eye(8), atol=1e-8)


def test_oracle_matches_live_case_A_complex(live):
    m = orc.solve(orc.make_model(live["A/left"].copy(), live["A/right"].copy()), complexify=True)
    np.testing.assert_allclose(m.sigma, live["A/cplx/sigma"], rtol=2e-5, atol=1e-5)
    orc.rotate(m, 6, 1)
    _check_rotated(live, "A/cplx/varimax", m, 6)


def test_oracle_matches_live_case_B_promax_f64(live):
    m = orc.solve(orc.make_model(live["B/left"].copy(), live["B/right"].copy()))
    assert m.rank == 40
    np.testing.assert_allclose(m.sigma[:39], live["B/sigma"][:39], rtol=1e-10)
    assert m.sigma[39] < 1e-12 * m.sigma[0]          # centring removes one dof
    orc.rotate(m, 8, 4)
    _check_rotated(live, "B/promax4", m, 8)


def test_oracle_matches_live_case_C_pca(live):
    m = orc.solve(orc.make_model(live["C/left"].copy()))
    np.testing.assert_allclose(m.sigma, live["C/sigma"], rtol=2e-5, atol=1e-5)
    orc.rotate(m, 5, 1)
    _check_rotated(live, "C/varimax", m, 5)


def test_oracle_rule_n_replays_reference_stream(live):
    m = orc.solve(orc.make_model(live["A/left"].copy(), live["A/right"].copy()))
    np.random.seed(123)
    got = orc.rule_n(m, 4, 10)
    assert got.shape == (10, 4)
    np.testing.assert_allclose(got, live["A/rule_n"], rtol=1e-8)


def test_mode_slice_semantics(live):
    m = orc.solve(orc.make_model(live["C/left"].copy()))
    assert orc.mode_slice(m, 3) == slice(0, 3)
    assert orc.mode_slice(m, None) == slice(0, m.rank)
    assert orc.mode_slice(m, slice(2, 4)) == slice(1, 4, None)
    assert orc.mode_slice(m, slice(None, 10 ** 6)) == slice(0, m.rank, None)
    with pytest.raises(ValueError):
        orc.mode_slice(m, 2.5)


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference (build container)")
def test_oracle_against_live_reference_fresh_inputs():
    from oracle.ref_harness import import_reference_mca
    MCA = import_reference_mca()
    A, B = orc.synthetic_fields(128, 200, 150, seed=3, k=12, dtype=np.float64)
    ref = MCA(A.copy(), B.copy()); ref.solve(); ref.rotate(10, 2)
    m = orc.rotate(orc.solve(orc.make_model(A.copy(), B.copy())), 10, 2)
    np.testing.assert_allclose(m.sigma[:100], ref.singular_values(100), rtol=1e-10)
    np.testing.assert_allclose(orc.get_variance(m), ref.variance(), rtol=1e-8)
    ge, re = orc.eofs(m, 10), ref.eofs(10)
    al, ar = orc.align_modes(re["left"], ge["left"], ge["right"])
    np.testing.assert_allclose(al, re["left"], atol=1e-7)
    np.testing.assert_allclose(ar, re["right"], atol=1e-7)
