"""Rule N (Overland & Preisendorfer 1982) on the GPU(s).

Semantics of xmca/array.py:1716-1771: ``n_runs`` times {two Gaussian fields of
the FULL grid size (NaN columns included), always float64 -> centre -> solve
[-> rotate] -> variance}; each spectrum is rescaled so that its sum equals the
sum of the model's own variances; rotated runs that do not converge are dropped.

B200 design: the runs are independent, so run indices are block-partitioned
over the ranks of a ``torch.distributed`` group (one process per GPU).  The
surrogates come from a counter-based Philox4x32-10 generator keyed by
(seed, global run index, field) -- the result does not depend on the number of
GPUs.  No collective is needed until the end: ONE all-gather of the
(modes x n_local) fp64 spectra (NCCL over NVLink; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L


def partition(n_runs: int, world: int, rank: int):
    """Contiguous block of run indices owned by ``rank`` (sizes differ by <= 1)."""
    base, extra = divmod(n_runs, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def _complex_solve(fields, extend, period):
    """Complex solve of centred real device fields: frequency-domain route, or the time-domain route on the
    exponentially extended series (array.py:455-472) when extend == 'exp'."""
    from . import device as D
    from . import engine as E
    A = fields[0]
    B = fields[1] if len(fields) > 1 else None
    if extend == "exp":
        H = D.hilbert_matrix_exp_extension(A.shape[0], float(period), A.dtype)
        Y = []
        for X in fields:
            y, _ = D.apply_time_operator(H, X)
            D.center_columns(y)
            Y.append(y)
        return E.solve_complex_time(A, Y[0], B, Y[1] if B is not None else None)
    return E.solve_complex(A, B)


def variance_of_fields(fields, complexify, rotated, n_rot, power, extend=False, period=1):
    """solve [+ rotate] + `_get_variance()` (sorted, array.py:771-779) of CENTRED real device fields
    (one or two, T x S_k): the body of one Monte-Carlo run (array.py:1757-1764, :1935-1945).
    Returns the variance spectrum (fp64 numpy, descending) or None if the rotation did not converge."""
    from . import device as D
    from . import engine as E
    t = D.torch()
    A = fields[0]
    B = fields[1] if len(fields) > 1 else None
    if not rotated:
        if complexify:
            sigma, _, _ = _complex_solve(fields, extend, period)
            return sigma
        return E.solve_real(A, B, want_vectors=False).sigma
    if complexify:
        sigma, vec, _ = _complex_solve(fields, extend, period)
        return _rotated_variance_complex(sigma, vec, len(fields), n_rot, power)
    return _rotated_variance(E.solve_real(A, B, want_vectors=True), B is None, n_rot, power)


def _rotated_variance(res, pca, n_rot, power):
    """rotate + sorted variance (array.py:823-834, :771-779) from a real solve result; None if not converged."""
    from . import device as D
    from . import engine as E
    t = D.torch()
    p = min(n_rot, res.sigma.size)
    root = D.to_device(np.sqrt(res.sigma[:p]))
    Vp = res.vectors(p)
    parts = [D.scale_copy(Vp[k], col_scale=root) for k in Vp]
    s_left = parts[0].shape[0]
    Ld = t.cat(parts, dim=0).contiguous() if len(parts) > 1 else parts[0]
    try:
        Lrot, _, _, _ = E.promax(Ld, power, max_iter=1000, tol=1e-8)
    except L.NotConvergedError:
        return None                                             # array.py:1759-1763
    nl = np.sqrt(D.to_host(D.col_sumsq(Lrot, 0, s_left)))
    var = nl ** 2 if pca else nl * np.sqrt(D.to_host(D.col_sumsq(Lrot, s_left, Lrot.shape[0])))
    return np.sort(var)[::-1]


def _surrogate_fields(shape_T, n_vars, run_index, seed, dtype):
    from . import device as D
    t = D.torch()
    fields = []
    for f, S in enumerate(n_vars):
        X = D.empty((shape_T, S), t.float32 if dtype == "float32" else t.float64)
        D.fill_normal(X, seed, 2 * run_index + f)               # array.py:1756
        D.center_columns(X)                                     # MCA ctor, array.py:199-207
        fields.append(X)
    return fields


def device_surrogate_variance(shape_T, n_vars, run_index, seed, complexify, rotated, n_rot, power,
                              dtype="float64"):
    """One surrogate run on the current CUDA device.  Returns the variance
    spectrum (fp64 numpy) or None if the rotation did not converge.
    dtype: storage precision of the Gaussian surrogate fields (see `rule_n`)."""
    fields = _surrogate_fields(shape_T, n_vars, run_index, seed, dtype)
    return variance_of_fields(fields, complexify, rotated, n_rot, power)


def _rotated_variance_complex(sigma, vec, n_fields, n_rot, power):
    from . import engine as E
    p = min(n_rot, sigma.size)
    keys = ["left", "right"][:n_fields]
    try:
        Br, Bi, s_left, _, _, _ = E.rotate_complex(vec.vectors(p), sigma, keys, p, power)
    except L.NotConvergedError:
        return None
    nl = E.complex_col_norms(Br, Bi, 0, s_left)
    var = nl ** 2 if n_fields == 1 else nl * E.complex_col_norms(Br, Bi, s_left, Br.shape[0])
    return np.sort(var)[::-1]


def device_surrogate_variance_pair(shape_T, n_vars, run_a, run_b, seed, rotated, n_rot, power, dtype="float64",
                                   complexify=False):
    """Two surrogate runs at once: the runs are independent (array.py:1753-1765), so their two symmetric
    eigenproblems go through ONE batched tridiagonalisation (engine.solve_real_pair).  Same Philox streams as
    the single-run path, hence the same surrogates.  Returns the two spectra (None where not converged)."""
    from . import engine as E
    fa = _surrogate_fields(shape_T, n_vars, run_a, seed, dtype)
    fb = _surrogate_fields(shape_T, n_vars, run_b, seed, dtype)
    pca = len(n_vars) == 1
    if complexify:
        (sa, va, _), (sb, vb, _) = E.solve_complex_pair(fa[0], None if pca else fa[1], fb[0], None if pca else fb[1],
                                                        want_vectors=rotated)
        del fa, fb
        if not rotated:
            return sa, sb
        return (_rotated_variance_complex(sa, va, len(n_vars), n_rot, power),
                _rotated_variance_complex(sb, vb, len(n_vars), n_rot, power))
    ra, rb = E.solve_real_pair(fa[0], None if pca else fa[1], fb[0], None if pca else fb[1], want_vectors=rotated)
    del fa, fb
    if not rotated:
        return ra.sigma, rb.sigma
    return _rotated_variance(ra, pca, n_rot, power), _rotated_variance(rb, pca, n_rot, power)


def gather_spectra(local: np.ndarray, valid: np.ndarray, n_runs: int, group=None):
    """All-gather the per-rank (modes x n_local) spectra into (modes x n_runs),
    ordered by global run index, dropping invalid columns."""
    import torch
    import torch.distributed as dist
    if group is None and not (dist.is_available() and dist.is_initialized()):
        return local[:, valid]
    world = dist.get_world_size(group)
    if world == 1:
        return local[:, valid]
    backend = dist.get_backend(group)
    devc = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    n_max = -(-n_runs // world)
    modes = local.shape[0]
    buf = torch.zeros((n_max, modes), dtype=torch.float64)
    buf[:local.shape[1]] = torch.from_numpy(np.nan_to_num(np.ascontiguousarray(local.T)))
    mask = torch.zeros(n_max, dtype=torch.float64)               # validity flag per run slot
    mask[:local.shape[1]] = torch.from_numpy(valid.astype(np.float64))
    send = torch.cat([buf.reshape(-1), mask]).to(devc)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)                    # the ONE collective of rule_n
    cols = []
    for r, chunk in enumerate(recv):
        chunk = chunk.cpu()
        spec = chunk[:n_max * modes].reshape(n_max, modes).numpy()
        ok = chunk[n_max * modes:].numpy() > 0.5
        n_r = len(partition(n_runs, world, r))
        cols.append(spec[:n_r][ok[:n_r]].T)
    return np.concatenate(cols, axis=1) if cols else np.zeros((modes, 0))


def _broadcast_seed(seed, group):
    """One seed for all ranks: rank 0's value (drawn from ITS global numpy stream when seed is None, the analogue of
    the reference consuming `np.random`, array.py:1756) is broadcast, so the Philox keys -- and therefore the result --
    do not depend on the number of GPUs."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else int(seed)
    rank = dist.get_rank(group)
    val = int(np.random.randint(0, 2 ** 31 - 1)) if (seed is None and rank == 0) else int(seed or 0)
    backend = dist.get_backend(group)
    devc = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([val], dtype=torch.int64, device=devc)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return int(t.item())


def rule_n(model, n_runs, n_modes=None, seed=None, group=None, _surrogate_fn=None, surrogate_dtype=None,
           pair_runs=True, _pair_fn=None):
    """Sharded Rule N for an ``xmca_b200.MCA`` model (array.py:1716-1771).

    surrogate_dtype: "float64" (default: the reference draws float64 surrogates whatever the model's dtype,
    array.py:1756) or "float32" -- an opt-in fast mode for fp32 models (fp32 Gaussian fields, Gram matrices on the
    tensor cores), statistically equivalent.  The returned spectra are fp64.
    seed: Philox key; None draws one from rank 0's global numpy stream and broadcasts it.
    pair_runs: the runs are processed two at a time (`device_surrogate_variance_pair`; same surrogates, same
    spectra to rounding)."""
    import torch.distributed as dist
    T = model._n_observations["left"]
    n_vars = [model._n_variables[k] for k in model._keys]
    complexify = model._analysis["is_complex"]
    rotated = model._analysis["is_rotated"]
    n_rot, power = model._analysis["n_rot"], model._analysis["power"]
    seed = _broadcast_seed(seed, group)
    if dist.is_available() and dist.is_initialized():
        world, rank = dist.get_world_size(group), dist.get_rank(group)
    else:
        world, rank = 1, 0
    fn = _surrogate_fn or device_surrogate_variance
    if surrogate_dtype is None:
        surrogate_dtype = "float64"
    if surrogate_dtype not in ("float32", "float64"):
        raise ValueError("surrogate_dtype must be 'float32' or 'float64'")
    extra = {} if _surrogate_fn is not None else {"dtype": surrogate_dtype}
    ref = model._get_variance()
    mine = partition(n_runs, world, rank)
    # The surrogates live on the FULL grid (NaN columns included, array.py:1745), so their spectra have
    # min(T, *n_vars) entries (n_rot when rotated) -- possibly more than the model's own rank when NaN columns were
    # dropped.  The reference stacks the full spectra and slices afterwards (array.py:1767-1771).
    modes = int(min(n_rot, T, *n_vars)) if rotated else int(min(T, *n_vars))
    local = np.full((modes, len(mine)), np.nan)
    valid = np.zeros(len(mine), dtype=bool)

    def store(j, spec):
        if spec is None:
            return
        spec = np.asarray(spec, dtype=np.float64)
        k = min(spec.size, modes)
        local[:, j] = 0.0
        local[:k, j] = spec[:k] * (ref.sum() / spec.sum())      # column-wise rescale, array.py:1768-1769
        valid[j] = True

    runs = list(mine)
    j = 0
    pairs = (_surrogate_fn is None or _pair_fn is not None) and pair_runs
    pair_fn = _pair_fn or device_surrogate_variance_pair
    while j < len(runs):
        if pairs and j + 1 < len(runs):
            sa, sb = pair_fn(T, n_vars, runs[j], runs[j + 1], seed, rotated, n_rot, power,
                             dtype=surrogate_dtype, complexify=complexify)
            store(j, sa)
            store(j + 1, sb)
            j += 2
        else:
            store(j, fn(T, n_vars, runs[j], seed, complexify, rotated, n_rot, power, **extra))
            j += 1
    sv = gather_spectra(local, valid, n_runs, group)
    return sv[model._get_slice(n_modes)]                    # array.py:1771 (bounded by the MODEL's rank, as there)
