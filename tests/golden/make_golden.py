"""Generate the committed golden vectors under tests/golden/.

Run in the BUILD CONTAINER (needs /root/reference):

    python tests/golden/make_golden.py

Outputs
-------
fixtures.npz    the reference's own golden fixtures re-packed (inputs sst/prcp,
                singular values std/cplx, first 100 (std) / 40 (cplx) EOFs)
live_cases.npz  outputs of the live, unmodified reference ``xmca.array.MCA``
                on small seeded inputs (the inputs are stored too): solve,
                rotate (Varimax / Promax), pcs, eofs, complex solve, PCA, rule_n
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.ref_harness import import_reference_mca, read_reference_fixtures  # noqa: E402


def planted(T, shape_l, shape_r, seed, k=8, dtype=np.float32, nan_cols=0):
    rng = np.random.default_rng(seed)
    sl, sr = int(np.prod(shape_l)), int(np.prod(shape_r))
    ts = rng.standard_normal((T, k))
    amp = 4.0 * 0.75 ** np.arange(k)
    A = (ts * amp) @ rng.standard_normal((k, sl)) / np.sqrt(k) + rng.standard_normal((T, sl))
    B = (ts * amp) @ rng.standard_normal((k, sr)) / np.sqrt(k) + rng.standard_normal((T, sr))
    A += 3.0 + rng.standard_normal(sl)       # non-zero means: exercises centring
    B -= 1.5
    A = A.astype(dtype)
    B = B.astype(dtype)
    if nan_cols:
        A[:, rng.choice(sl, nan_cols, replace=False)] = np.nan
    return A.reshape((T,) + shape_l), B.reshape((T,) + shape_r)


def dump_state(out, tag, model, n):
    out[tag + "/sigma"] = model.singular_values()
    V = model._get_V(model._analysis["rank"], rotated=False)
    for k, v in V.items():
        out[tag + "/V_" + k] = v
    for k, v in model.pcs(n, rotated=False).items():
        out[tag + "/pcs_unrot_" + k] = v
    for k, v in model.eofs(n, rotated=False).items():
        out[tag + "/eofs_unrot_" + k] = v
    out[tag + "/explained_variance"] = model.explained_variance(n)
    out[tag + "/total"] = np.array([model._analysis["total_covariance"],
                                    model._analysis["total_squared_covariance"]], dtype=np.float64)


def dump_rotated(out, tag, model, n):
    out[tag + "/R"] = model.rotation_matrix()
    out[tag + "/Rit"] = model.rotation_matrix(inverse_transpose=True)
    out[tag + "/Phi"] = model.correlation_matrix()
    out[tag + "/variance"] = model.variance(n)
    out[tag + "/var_idx"] = model._var_idx
    for k, v in model.norm(n).items():
        out[tag + "/norm_" + k] = v
    for k, v in model.pcs(n).items():
        out[tag + "/pcs_" + k] = v
    for k, v in model.eofs(n).items():
        out[tag + "/eofs_" + k] = v
    for k, v in model.eofs(slice(2, 4), scaling="max").items():
        out[tag + "/eofs_slice_max_" + k] = v
    for k, v in model.pcs(3, scaling="std").items():
        out[tag + "/pcs_std_" + k] = v


def main():
    MCA = import_reference_mca()
    fx = read_reference_fixtures()
    np.savez_compressed(
        os.path.join(HERE, "fixtures.npz"),
        sst=fx["sst"], prcp=fx["prcp"], sv_std=fx["sv_std"], sv_cplx=fx["sv_cplx"],
        eofs_std_sst=fx["eofs_std_sst"][..., :100], eofs_std_prcp=fx["eofs_std_prcp"][..., :100],
        eofs_cplx_sst=fx["eofs_cplx_sst"][..., :40].astype(np.complex128),
        eofs_cplx_prcp=fx["eofs_cplx_prcp"][..., :40].astype(np.complex128))

    out = {}
    # case A: T > S (direct route), f32, NaN columns, 2-D space on the left
    A, B = planted(96, (6, 10), (50,), seed=11, nan_cols=3)
    out["A/left"], out["A/right"] = A, B
    m = MCA(A.copy(), B.copy()); m.solve()
    dump_state(out, "A", m, 8)
    m.rotate(8, 1); dump_rotated(out, "A/varimax", m, 8)
    m = MCA(A.copy(), B.copy()); m.solve(); m.rotate(8, 2)
    dump_rotated(out, "A/promax2", m, 8)
    m = MCA(A.copy(), B.copy()); m.solve(complexify=True)
    dump_state(out, "A/cplx", m, 6)
    m.rotate(6, 1); dump_rotated(out, "A/cplx/varimax", m, 6)
    # rule_n with the global numpy RNG stream (array.py:1756)
    m = MCA(A.copy(), B.copy()); m.solve()
    np.random.seed(123)
    out["A/rule_n"] = m.rule_n(4, 10)

    # case B: T < S (Gram route; centring leaves one null mode), f64
    A, B = planted(40, (96,), (8, 10), seed=12, dtype=np.float64)
    out["B/left"], out["B/right"] = A, B
    m = MCA(A.copy(), B.copy()); m.solve()
    dump_state(out, "B", m, 8)
    m.rotate(8, 4); dump_rotated(out, "B/promax4", m, 8)

    # case C: PCA of a single field
    A, _ = planted(60, (5, 9), (4,), seed=13)
    out["C/left"] = A
    m = MCA(A.copy()); m.solve()
    dump_state(out, "C", m, 5)
    m.rotate(5, 1); dump_rotated(out, "C/varimax", m, 5)

    np.savez_compressed(os.path.join(HERE, "live_cases.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
