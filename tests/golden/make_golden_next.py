"""Golden vectors for the callers either side of the hot path: predict,
reconstructed_fields, homogeneous / heterogeneous patterns and bootstrapping, produced by the live,
unmodified reference ``xmca.array.MCA`` in the BUILD CONTAINER (needs /root/reference):

    python tests/golden/make_golden_next.py     ->  tests/golden/live_next.npz

bootstrapping uses the global numpy RNG (tools/array.py:132): the seed set before each call is
recorded so that the engine can replay the identical index stream.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.ref_harness import import_reference_mca  # noqa: E402
from make_golden import planted  # noqa: E402


def dump(out, tag, m, A, B, n):
    new = {"left": A[:20].copy()}
    if B is not None:
        new["right"] = B[:20].copy()
    for scaling in ("None", "eigen", "max"):
        for k, v in m.predict(n=n, scaling=scaling, **new).items():
            out["%s/predict_%s_%s" % (tag, scaling, k)] = v
    for k, v in m.reconstructed_fields(mode=3, original_scale=True).items():
        out["%s/rec3_orig_%s" % (tag, k)] = v
    for k, v in m.reconstructed_fields(mode=slice(2, 4), original_scale=False).items():
        out["%s/rec24_%s" % (tag, k)] = v
    r, p = m.homogeneous_patterns(n)
    for k in r:
        out["%s/hom_r_%s" % (tag, k)] = r[k]
        out["%s/hom_p_%s" % (tag, k)] = p[k]
    if B is not None:
        r, p = m.heterogeneous_patterns(n)
        for k in r:
            out["%s/het_r_%s" % (tag, k)] = r[k]
            out["%s/het_p_%s" % (tag, k)] = p[k]


def main():
    MCA = import_reference_mca()
    out = {}
    A, B = planted(96, (6, 10), (50,), seed=11, nan_cols=3)
    out["A/left"], out["A/right"] = A, B
    m = MCA(A.copy(), B.copy()); m.solve()
    dump(out, "A", m, A, B, 5)
    m.rotate(6, 1)
    dump(out, "A/varimax", m, A, B, 5)
    m = MCA(A.copy(), B.copy()); m.solve(); m.rotate(6, 2)
    dump(out, "A/promax2", m, A, B, 5)
    m = MCA(A.copy(), B.copy()); m.normalize(); m.solve()
    dump(out, "A/normalized", m, A, B, 5)
    m = MCA(A.copy(), B.copy()); m.solve(complexify=True)
    dump(out, "A/cplx", m, A, B, 4)
    # complex solve on the exponentially extended series (array.py:378-411, :455-472)
    m = MCA(A.copy(), B.copy()); m.solve(complexify=True, extend="exp", period=12)
    out["A/cplx_exp/sigma"] = m.singular_values()
    out["A/cplx_exp/field_left"] = m._fields["left"]
    for k, v in m.eofs(4).items():
        out["A/cplx_exp/eofs_" + k] = v
    for k, v in m.pcs(4).items():
        out["A/cplx_exp/pcs_" + k] = v
    # PCA
    C, _ = planted(60, (5, 9), (4,), seed=13)
    out["C/left"] = C
    m = MCA(C.copy()); m.solve()
    dump(out, "C", m, C, None, 4)

    # bootstrapping with the recorded global seed
    m = MCA(A.copy(), B.copy()); m.solve()
    np.random.seed(77)
    out["A/boot_time_left"] = m.bootstrapping(5, n_modes=6, on_left=True, on_right=False, disable_progress=True)
    np.random.seed(78)
    out["A/boot_time_both_block4"] = m.bootstrapping(4, n_modes=6, on_left=True, on_right=True, block_size=4,
                                                     disable_progress=True)
    np.random.seed(79)
    out["A/boot_space_perm"] = m.bootstrapping(3, n_modes=5, axis=1, on_left=False, on_right=True, replace=False,
                                               disable_progress=True)
    np.random.seed(80)
    out["A/boot_iterative"] = m.bootstrapping(3, n_modes=3, strategy="iterative", disable_progress=True)
    m.rotate(5, 1)
    np.random.seed(81)
    out["A/varimax/boot"] = m.bootstrapping(4, n_modes=4, on_left=True, on_right=True, disable_progress=True)
    m = MCA(A.copy(), B.copy()); m.solve(complexify=True, extend="exp", period=12)
    np.random.seed(82)
    out["A/cplx_exp/boot"] = m.bootstrapping(3, n_modes=4, on_left=True, on_right=True, block_size=2,
                                             disable_progress=True)
    m = MCA(A.copy(), B.copy()); m.solve(complexify=True)
    np.random.seed(83)
    out["A/cplx/boot"] = m.bootstrapping(3, n_modes=4, on_left=True, on_right=False, disable_progress=True)
    np.savez_compressed(os.path.join(HERE, "live_next.npz"), **out)
    print("wrote live_next.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
