"""CPU tests: host-side class logic, C-ABI export surface, rule_n sharding over
a 2-process gloo group.  No GPU compute is attempted here."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import mca_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    from xmca_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    return _lib


def test_library_exports_every_declared_symbol():
    _lib = _ensure_built()
    lib = _lib.load()
    names = _lib.declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert lib.xmca_version() >= 100
    assert lib.xmca_jacobi_padded_cols(155) == 192 and lib.xmca_jacobi_padded_cols(1) == 64
    assert lib.xmca_gemm_workspace_bytes(10, 10, 4, 1) == 10 * 10 * 4 * 8
    assert lib.xmca_jacobi_workspace_bytes(492, 155) > 0
    assert lib.xmca_varimax_workspace_bytes(1000, 20) > 1000 * 20 * 8


def test_bad_arguments_are_rejected_without_a_gpu():
    _lib = _ensure_built()
    lib = _lib.load()
    rc = lib.xmca_gemm(1, 1, 0, 4, 4, 1.0, None, 0, 4, None, 0, 4, None, 0, 4, 0, 1, 1, None, 0, None)
    assert rc == _lib.BAD_ARG
    assert b"xmca_gemm" in lib.xmca_last_error()
    rc = lib.xmca_varimax(None, 0, 10, 80, 80, 1.0, 10, 1e-8, None, 80, None, None, None, None, 0, None)
    assert rc == _lib.BAD_ARG


def test_constructor_validation_matches_reference_unit_tests():
    from xmca_b200 import MCA
    rng = np.random.default_rng(7)
    A = rng.standard_normal((500, 20, 15))
    B = rng.standard_normal((500, 15, 10))
    MCA()
    MCA(A)
    m = MCA(A, B)
    assert m._analysis["method"] == "mca" and m._analysis["is_bivariate"]
    assert m._fields["left"].shape == (500, 300) and abs(m._fields["left"].mean(axis=0)).max() < 1e-12
    with pytest.raises(ValueError):
        MCA(A, B, A)
    with pytest.raises(ValueError):
        MCA(A[:20], B[:15])
    with pytest.raises(TypeError):
        MCA(list(A))
    bad = A.copy()
    bad[3] = np.nan
    with pytest.raises(ValueError):
        MCA(bad)
    hole = A.copy()
    hole[:, 2, 3] = np.nan                       # a NaN grid point is dropped, not an error
    m = MCA(hole)
    assert m._fields["left"].shape == (500, 299) and m._n_variables["left"] == 300
    assert m.fields()["left"].shape == A.shape and np.isnan(m.fields()["left"][:, 2, 3]).all()
    np.testing.assert_allclose(np.nan_to_num(m.fields(original_scale=True)["left"]), np.nan_to_num(hole), atol=1e-12)


def test_getters_before_solve_raise_runtime_error():
    from xmca_b200 import MCA
    m = MCA(np.random.default_rng(0).standard_normal((30, 5)))
    for call in (m.singular_values, m.eofs, m.pcs, m.norm, m.variance):
        with pytest.raises(RuntimeError):
            call()
    with pytest.raises(RuntimeError):
        MCA().solve()


def test_mode_slice_semantics_match_oracle():
    from xmca_b200 import MCA
    m = MCA(np.random.default_rng(0).standard_normal((30, 5)))
    m._analysis["rank"] = 5
    ref = orc.solve(orc.make_model(np.random.default_rng(0).standard_normal((30, 5))))
    for n in (None, 3, slice(2, 4), slice(None, 99), slice(1, None)):
        assert m._get_slice(n) == orc.mode_slice(ref, n)
    with pytest.raises(ValueError):
        m._get_slice(1.5)


def test_solve_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from xmca_b200 import MCA
    from xmca_b200._lib import XmcaLibraryError
    m = MCA(np.random.default_rng(0).standard_normal((30, 5)))
    with pytest.raises(XmcaLibraryError):
        m.solve()


def test_rule_n_partition_covers_all_runs():
    from xmca_b200.rule_n import partition
    for n, w in [(10, 1), (10, 3), (1000, 8), (3, 8)]:
        got = [i for r in range(w) for i in partition(n, w, r)]
        assert got == list(range(n))
        sizes = [len(partition(n, w, r)) for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rule_n_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from xmca_b200 import MCA
    from xmca_b200.rule_n import rule_n
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    m = MCA(rng.standard_normal((40, 12)), rng.standard_normal((40, 9)))
    # stand-in for a solved model: only the fields rule_n reads
    m._analysis["rank"] = 9
    m._norm = {"left": np.sqrt(np.linspace(9, 1, 9)), "right": np.sqrt(np.linspace(9, 1, 9))}
    m._var_idx = np.arange(9)
    m._singular_values = np.linspace(9, 1, 9)

    def stub(T, n_vars, run, seed, cplx, rot, n_rot, power):          # deterministic in the run index
        if run == 5:
            return None                                               # a dropped (non-converged) run
        r = np.random.default_rng(1000 * seed + run)
        return np.sort(r.random(9))[::-1] + run

    got = rule_n(m, 11, n_modes=4, seed=17, _surrogate_fn=stub)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), got)

    # the same runs two at a time (what the GPU path does): identical bookkeeping, also with an odd number of
    # runs on a rank and a dropped run inside a pair
    calls = []

    def pair_stub(T, n_vars, run_a, run_b, seed, rot, n_rot, power, dtype=None, complexify=False):
        calls.append((run_a, run_b))
        return (stub(T, n_vars, run_a, seed, complexify, rot, n_rot, power),
                stub(T, n_vars, run_b, seed, complexify, rot, n_rot, power))

    paired = rule_n(m, 11, n_modes=4, seed=17, _surrogate_fn=stub, _pair_fn=pair_stub)
    np.testing.assert_array_equal(paired, got)
    mine = list(__import__("xmca_b200.rule_n", fromlist=["partition"]).partition(11, world, rank))
    assert calls == [(mine[k], mine[k + 1]) for k in range(0, len(mine) - 1, 2)]

    # seed=None: ONE seed, drawn on rank 0 and broadcast -- every rank must use the same Philox key whatever its own
    # numpy state is (the ranks are given different states on purpose)
    seen = []

    def seed_stub(T, n_vars, run, seed, cplx, rot, n_rot, power):
        seen.append(seed)
        return stub(T, n_vars, run, seed % 1000, cplx, rot, n_rot, power)

    np.random.seed(5 if rank == 0 else 999)
    none_seed = rule_n(m, 7, n_modes=4, seed=None, _surrogate_fn=seed_stub)
    assert len(set(seen)) == 1
    np.save(os.path.join(out_dir, "seed%d.npy" % rank), np.array(seen[:1]))
    np.save(os.path.join(out_dir, "none%d.npy" % rank), none_seed)

    # NaN columns: the model's rank (7 here) is smaller than the length of the surrogates' spectra (9): the
    # reference stacks the full spectra, rescales by their FULL sum and slices by the model's rank (array.py:1767-1771)
    m2 = MCA(rng.standard_normal((40, 12)), rng.standard_normal((40, 9)))
    m2._analysis["rank"] = 7
    m2._norm = {"left": np.sqrt(np.linspace(7, 1, 7)), "right": np.sqrt(np.linspace(7, 1, 7))}
    m2._var_idx = np.arange(7)
    m2._singular_values = np.linspace(7, 1, 7)
    nan_case = rule_n(m2, 4, seed=3, _surrogate_fn=stub)
    np.save(os.path.join(out_dir, "nan%d.npy" % rank), nan_case)
    dist.destroy_process_group()


def test_rule_n_two_rank_gloo_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_rule_n_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    np.testing.assert_array_equal(r0, r1)
    assert r0.shape == (4, 10)                                        # 11 runs, one dropped
    # single-process result with the same stub
    cols = []
    ref_sum = np.linspace(9, 1, 9).sum()
    for run in range(11):
        if run == 5:
            continue
        s = np.sort(np.random.default_rng(1000 * 17 + run).random(9))[::-1] + run
        cols.append(s * ref_sum / s.sum())
    np.testing.assert_allclose(r0, np.array(cols).T[:4], rtol=1e-14)
    # seed=None: both ranks used rank 0's draw; the result equals a single-process run with the same numpy state
    s0, s1 = np.load(tmp_path / "seed0.npy"), np.load(tmp_path / "seed1.npy")
    assert s0[0] == s1[0]
    np.random.seed(5)
    assert s0[0] == np.random.randint(0, 2 ** 31 - 1)
    np.testing.assert_array_equal(np.load(tmp_path / "none0.npy"), np.load(tmp_path / "none1.npy"))
    # NaN-column case: (rank 7 rows) x 4 runs, columns rescaled by the full 9-entry sum
    nan0 = np.load(tmp_path / "nan0.npy")
    assert nan0.shape == (7, 4)
    want = []
    for run in range(4):
        sp = np.sort(np.random.default_rng(1000 * 3 + run).random(9))[::-1] + run
        want.append(sp * np.linspace(7, 1, 7).sum() / sp.sum())
    np.testing.assert_allclose(nan0, np.array(want).T[:7], rtol=1e-14)


def test_xmca_facade_constructor_and_metadata():
    """xmca/xarray.py:31-85 and tests/unit/test_xarray.py:30-38: DataArrays only, <= 2 fields,
    dims/coords kept; works with a duck-typed backend because xarray is absent from the image."""
    import xr_stub as stub
    from xmca_b200 import xarray as X
    X.set_backend(stub)
    rng = np.random.default_rng(3)
    t, lat, lon = np.arange(30), np.linspace(-60, 60, 5), np.linspace(0, 300, 6)
    da = stub.DataArray(rng.standard_normal((30, 5, 6)), dims=("time", "lat", "lon"),
                        coords={"time": t, "lat": lat, "lon": lon}, name="sst")
    m = X.xMCA(da, da)
    assert m._field_dims["left"] == ("time", "lat", "lon") and m._n_variables["right"] == 30
    f = m.fields()
    assert isinstance(f["left"], stub.DataArray) and f["left"].dims == ("time", "lat", "lon")
    with pytest.raises(TypeError):
        X.xMCA(np.zeros((30, 5, 6)))
    with pytest.raises(ValueError):
        X.xMCA(da, da, da)
    with pytest.raises(RuntimeError):
        m.singular_values()
    m.apply_coslat()
    assert m._analysis["is_coslat_corrected"]
    w = np.sqrt(np.cos(np.deg2rad(lat)) + 1e-6)
    want = (da.values - da.values.mean(axis=0)) * w[None, :, None]
    np.testing.assert_allclose(m.fields()["left"].values, want, atol=1e-12)


def test_info_file_round_trip(tmp_path):
    """info.xmca (array.py:1629-1714): same `key : value` layout, values re-typed from the defaults."""
    from xmca_b200 import MCA
    rng = np.random.default_rng(1)
    m = MCA(rng.standard_normal((20, 6)), rng.standard_normal((20, 4)))
    m.set_field_names("sea surface temp", "precip")
    m._analysis.update({"is_rotated": True, "n_rot": 7, "power": 2, "rank": 4, "is_complex": True,
                        "total_covariance": 12.5, "is_normalized": True})
    m._create_info_file(str(tmp_path))
    text = (tmp_path / "info.xmca").read_text()
    assert "\nleft                 : sea surface temp" in text and "\nn_rot                : 7" in text
    assert m._get_file_names("nc")["eofs"]["left"] == "sea_surface_temp_eofs.nc"
    m2 = MCA()
    m2._set_info_from_file(str(tmp_path / "info.xmca"))
    for k in ("is_rotated", "n_rot", "power", "rank", "is_complex", "is_bivariate", "is_normalized", "method"):
        assert m2._analysis[k] == m._analysis[k], k
    assert m2._analysis["total_covariance"] == 12.5 and m2._field_names["right"] == "precip"
