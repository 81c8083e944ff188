"""GPU parity of the callers downstream of the hot path against outputs of the
live reference (tests/golden/live_next.npz, generator tests/golden/make_golden_next.py)."""
import numpy as np
import pytest

from oracle import mca_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def MCA():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from xmca_b200 import MCA
    return MCA


def _model(MCA, g, tag):
    A, B = g["A/left"], g["A/right"]
    if tag == "C":
        m = MCA(g["C/left"].copy())
        m.solve()
        return m, g["C/left"], None
    m = MCA(A.copy(), B.copy())
    if tag == "A/normalized":
        m.normalize()
    m.solve(complexify=(tag == "A/cplx"))
    if tag == "A/varimax":
        m.rotate(6, 1)
    if tag == "A/promax2":
        m.rotate(6, 2)
    return m, A, B


def _aligned(ref, got):
    """Align the mode signs / phases of `got` (.., m) to `ref` and return it."""
    r2, g2 = np.nan_to_num(ref).reshape(-1, ref.shape[-1]), np.nan_to_num(got).reshape(-1, got.shape[-1])
    f = np.sum(np.conj(g2) * r2, axis=0)
    f = np.where(np.abs(f) > 0, f / np.maximum(np.abs(f), 1e-300), 1.0)
    return got * f


@pytest.mark.parametrize("tag,n", [("A", 5), ("A/varimax", 5), ("A/promax2", 5), ("A/normalized", 5),
                                   ("A/cplx", 4), ("C", 4)])
def test_predict_reconstruct_patterns_match_live_reference(MCA, live_next, tag, n):
    g = live_next
    m, A, B = _model(MCA, g, tag)
    new = {"left": A[:20].copy()}
    if B is not None:
        new["right"] = B[:20].copy()
    keys = list(new)
    # ("max" divides by max |Re pcs|, which depends on the arbitrary phase of a complex mode)
    for scaling in (("None", "eigen") if tag == "A/cplx" else ("None", "eigen", "max")):
        got = m.predict(n=n, scaling=scaling, **new)
        for k in keys:
            ref = g["%s/predict_%s_%s" % (tag, scaling, k)]
            assert got[k].shape == ref.shape
            # left and right flip together: align on the left field
            fac = _aligned(g["%s/predict_%s_left" % (tag, scaling)], got["left"]) / np.where(got["left"] == 0, 1, got["left"])
            np.testing.assert_allclose(got[k] * fac[0], ref, atol=3e-3 * np.abs(ref).max())
    rec = m.reconstructed_fields(mode=3, original_scale=True)
    rec24 = m.reconstructed_fields(mode=slice(2, 4), original_scale=False)
    for k in keys:
        ref = g["%s/rec3_orig_%s" % (tag, k)]
        assert rec[k].shape == ref.shape and np.array_equal(np.isnan(rec[k]), np.isnan(ref))
        np.testing.assert_allclose(np.nan_to_num(rec[k]), np.nan_to_num(ref), atol=2e-3 * np.nanmax(np.abs(ref)))
        ref = g["%s/rec24_%s" % (tag, k)]
        np.testing.assert_allclose(np.nan_to_num(rec24[k]), np.nan_to_num(ref), atol=2e-3 * np.nanmax(np.abs(ref)) + 1e-4)
    r, p = m.homogeneous_patterns(n)
    if tag == "A/cplx":
        # correlations with Re(pcs) depend on the arbitrary phase of each complex mode: shapes / ranges only
        for k in keys:
            assert r[k].shape == g["%s/hom_r_%s" % (tag, k)].shape
            assert np.nanmax(np.abs(r[k])) <= 1.0 and np.nanmin(p[k]) >= 0.0 and np.nanmax(p[k]) <= 1.0
        return
    for k in keys:
        rr, pr = g["%s/hom_r_%s" % (tag, k)], g["%s/hom_p_%s" % (tag, k)]
        assert r[k].shape == rr.shape
        np.testing.assert_allclose(np.abs(np.nan_to_num(r[k])), np.abs(np.nan_to_num(rr)), atol=2e-3)
        np.testing.assert_allclose(np.nan_to_num(p[k]), np.nan_to_num(pr), atol=2e-2)
    if B is not None:
        r, p = m.heterogeneous_patterns(n)
        for k in keys:
            rr = g["%s/het_r_%s" % (tag, k)]
            np.testing.assert_allclose(np.abs(np.nan_to_num(r[k])), np.abs(np.nan_to_num(rr)), atol=2e-3)
    else:
        with pytest.raises(KeyError):
            m.heterogeneous_patterns(n)


def test_predict_reproduces_pcs_and_rejects_bad_shapes(MCA, live_next):
    """test_integration_xarray.py:368-430: predict(first 20 steps) == pcs()[:20]; wrong dims -> ValueError."""
    g = live_next
    for rot in (None, (6, 1), (6, 3)):
        m = MCA(g["A/left"].copy(), g["A/right"].copy())
        m.solve()
        if rot:
            m.rotate(*rot)
        got = m.predict(g["A/left"][:20], g["A/right"][:20], n=4)
        pcs = m.pcs(4)
        for k in ("left", "right"):
            np.testing.assert_allclose(got[k], pcs[k][:20], atol=2e-4 * np.abs(pcs[k]).max())
    with pytest.raises(ValueError):
        m.predict(g["A/left"][:20, 0])
    with pytest.raises(ValueError):
        m.predict(g["A/left"][:20, :3])
    with pytest.raises(ValueError):
        m.predict(g["A/left"][:20], n=3, scaling="bogus")


@pytest.mark.parametrize("key,seed,kw", [
    ("A/boot_time_left", 77, dict(n_runs=5, n_modes=6, on_left=True, on_right=False)),
    ("A/boot_time_both_block4", 78, dict(n_runs=4, n_modes=6, on_left=True, on_right=True, block_size=4)),
    ("A/boot_space_perm", 79, dict(n_runs=3, n_modes=5, axis=1, on_left=False, on_right=True, replace=False)),
    ("A/boot_iterative", 80, dict(n_runs=3, n_modes=3, strategy="iterative")),
])
def test_bootstrapping_replays_the_reference_stream(MCA, live_next, key, seed, kw):
    """array.py:1813-1952 with the same global-RNG index stream as the live reference: identical spectra."""
    g = live_next
    m = MCA(g["A/left"].copy(), g["A/right"].copy())
    m.solve()
    np.random.seed(seed)
    got = m.bootstrapping(disable_progress=True, **kw)
    ref = g[key]
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=2e-4, atol=2e-5 * ref.max())


def test_bootstrapping_of_a_rotated_model(MCA, live_next):
    g = live_next
    m = MCA(g["A/left"].copy(), g["A/right"].copy())
    m.solve()
    m.rotate(5, 1)
    np.random.seed(81)
    got = m.bootstrapping(4, n_modes=4, on_left=True, on_right=True, disable_progress=True)
    ref = g["A/varimax/boot"]
    assert got.shape == ref.shape
    np.testing.assert_allclose(got, ref, rtol=2e-3, atol=1e-4 * ref.max())
    with pytest.raises(ValueError):
        m.bootstrapping(2, block_size=7, disable_progress=True)


@pytest.mark.parametrize("variant", ["plain", "rotated_normalized", "complex"])
def test_load_analysis_round_trip(MCA, live_next, tmp_path, variant):
    """test_integration_xarray.py:87-148 for the ndarray class: info.xmca + original-scale fields + unrotated
    EOFs + singular values rebuild the model (array.py:1954-2012), including the re-run rotation."""
    g = live_next
    A, B = g["A/left"], g["A/right"]
    m = MCA(A.copy(), B.copy())
    if variant == "rotated_normalized":
        m.normalize()
    m.solve(complexify=(variant == "complex"))
    if variant == "rotated_normalized":
        m.rotate(5, 2)
    m._create_info_file(str(tmp_path))
    fields = {k: np.real(v) for k, v in m.fields(original_scale=True).items()}
    eofs = m.eofs(rotated=False)
    sv = m.singular_values()
    m2 = MCA()
    m2.load_analysis(str(tmp_path / "info.xmca"), fields=fields, eofs=eofs, singular_values=sv)
    assert m2._analysis["rank"] == m._analysis["rank"] and m2._solve_info["route"] == "loaded"
    np.testing.assert_array_equal(m2.singular_values(), sv)
    e1, e2 = m.eofs(4), m2.eofs(4)
    p1, p2 = m.pcs(4), m2.pcs(4)
    for k in ("left", "right"):
        np.testing.assert_allclose(np.nan_to_num(e2[k]), np.nan_to_num(e1[k]), atol=1e-4 * np.nanmax(np.abs(e1[k])))
        np.testing.assert_allclose(p2[k], p1[k], atol=2e-4 * np.abs(p1[k]).max())
    np.testing.assert_allclose(m2.variance(4), m.variance(4), rtol=1e-4)


def test_complex_solve_with_exponential_extension(MCA, live_next):
    """solve(complexify=True, extend='exp', period) -- array.py:378-411, :455-472: the Hilbert transform of
    the fore/back-cast series as ONE T x T operator on the device, against the live reference."""
    g = live_next
    m = MCA(g["A/left"].copy(), g["A/right"].copy())
    m.solve(complexify=True, extend="exp", period=12)
    ref = g["A/cplx_exp/sigma"]
    np.testing.assert_allclose(m.singular_values(30), ref[:30], rtol=5e-5)
    np.testing.assert_allclose(m.singular_values(), ref, atol=2e-5 * ref[0])
    zr = g["A/cplx_exp/field_left"]
    np.testing.assert_allclose(m._fields["left"], zr, atol=2e-5 * np.abs(zr).max())
    e, p = m.eofs(4), m.pcs(4)
    er, pr = g["A/cplx_exp/eofs_left"], g["A/cplx_exp/pcs_left"]
    np.testing.assert_allclose(_aligned(er, e["left"]), er, atol=1e-3 * np.nanmax(np.abs(er)), equal_nan=True)
    np.testing.assert_allclose(_aligned(pr, p["left"]), pr, atol=1e-3 * np.abs(pr).max())
    with pytest.raises(NotImplementedError):
        m.solve(complexify=True, extend="theta", period=12)
    with pytest.raises(ValueError):
        m.solve(complexify=True, extend="bogus")
    m.solve(complexify=True)                                  # back to the plain transform: caches are dropped
    np.testing.assert_allclose(m.singular_values(10), orc.solve(orc.make_model(g["A/left"].copy(), g["A/right"].copy()),
                                                                  complexify=True).sigma[:10], rtol=5e-5)


@pytest.mark.parametrize("key,seed,ext,kw", [
    ("A/cplx_exp/boot", 82, "exp", dict(n_runs=3, n_modes=4, on_left=True, on_right=True, block_size=2)),
    ("A/cplx/boot", 83, False, dict(n_runs=3, n_modes=4, on_left=True, on_right=False)),
])
def test_bootstrapping_of_complex_models(MCA, live_next, key, seed, ext, kw):
    g = live_next
    m = MCA(g["A/left"].copy(), g["A/right"].copy())
    m.solve(complexify=True, extend=ext, period=12 if ext else 1)
    np.random.seed(seed)
    got = m.bootstrapping(disable_progress=True, **kw)
    np.testing.assert_allclose(got, g[key], rtol=5e-4, atol=5e-5 * g[key].max())
