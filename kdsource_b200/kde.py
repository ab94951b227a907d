"""Bandwidth selection on the GPU, mirroring ``python/kdsource/kde.py``.

Same public names, arguments and error behaviour as the reference module
(``bw_silv``, ``bw_knn``, ``_kde_cv_score``, ``bw_mlcv``, ``bw_auto``,
``optimize_bw``, ``bw_methods``); the arithmetic that the reference delegates
to scikit-learn (``NearestNeighbors``, ``KFold``), joblib and KDEpy runs in
the ``knn`` and ``mlcv_score`` CUDA kernels behind the C ABI.

Multi-GPU: when ``torch.distributed`` is initialised with more than one rank,
``bw_knn`` shards whole batches and ``mlcv_scores`` shards test rows over the
ranks (sources replicated); results are combined with an all-gather /
all-reduce and are identical on every rank.
"""

import ctypes

import numpy as np

from . import _lib
from . import dist as _dist
from .treekde import TreeKDE

_LOG2E = 1.4426950408889634


def _lib_is_tensor(x):
    """True for a torch tensor (without importing torch for plain numpy input)."""
    return type(x).__module__.startswith("torch")


def bw_silv(dim, N_eff):
    """Silverman's Rule (reference kde.py:18-38)."""
    return (4 / (2 + dim)) ** (1 / (4 + dim)) * N_eff ** (-1 / (4 + dim))


# ---------------------------------------------------------------------------
# kNN
# ---------------------------------------------------------------------------

def array_split_bounds(N, batches):
    """Boundaries of ``np.array_split(data, batches)`` (kde.py:91)."""
    sizes = np.full(batches, N // batches, dtype=np.int64)
    sizes[: N % batches] += 1
    return np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)


def knn_batched(data, kk, batch_off, precision="float64", want_dist=False, want_idx=False):
    """kk nearest neighbours (self included) of every row inside its batch.

    Returns (kth, dist, idx): ``kth`` (N,) distance to the kk-th neighbour,
    ``dist`` (N,kk) ascending, ``idx`` (N,kk) int32 indices inside the batch
    ordered by (distance, index).  Batches are sharded over ranks.  In a batch
    with fewer than kk rows the missing neighbours are (+inf, -1); ``bw_knn``
    refuses that case like scikit-learn does.  `data` may be a (N,D) float64 CUDA
    tensor (device-resident fit); ``kth`` is then returned as a CUDA tensor.
    """
    L = _lib.load()
    t = _lib.torch()
    on_device = t.is_tensor(data)  # (N,D) float64 CUDA tensor: stays where it is
    if on_device:
        data = data.to(dtype=t.float64).contiguous()
    else:
        data = np.ascontiguousarray(data, dtype=np.float64)
    N, D = data.shape
    batch_off = np.ascontiguousarray(batch_off, dtype=np.int64)
    nb = len(batch_off) - 1
    if nb < 1 or N == 0:
        raise ValueError("knn_batched needs at least one non-empty batch")
    rank, world = _dist.rank_world()
    b_lo, b_hi = _dist.shard_range(nb, rank, world)
    r_lo, r_hi = int(batch_off[b_lo]), int(batch_off[b_hi])
    n_loc = r_hi - r_lo
    t_kth = _lib.empty(n_loc, t.float64)
    t_dist = _lib.empty(n_loc * kk, t.float64) if want_dist else None
    t_idx = _lib.empty(n_loc * kk, t.int32) if want_idx else None
    use_f64 = precision == "float64"

    def run(d_rows, g_lo, g_hi, row0):
        """batches [g_lo, g_hi) of this rank; their rows start at local row `row0`"""
        offs = batch_off[g_lo:g_hi + 1] - batch_off[g_lo]
        n = int(offs[-1])
        d_off = _lib.to_dev(offs, np.int64)
        scratch = _lib.empty(n * D, t.float64 if use_f64 else t.float32)
        _lib.check(L.kdsb200_knn(
            _lib.ptr(d_rows), n, D, _lib.ptr(d_off), g_hi - g_lo, int(np.max(np.diff(offs))),
            kk, int(use_f64), _lib.ptr(scratch), _lib.ptr(t_kth[row0:]),
            _lib.ptr(t_dist[row0 * kk:]) if want_dist else None,
            _lib.ptr(t_idx[row0 * kk:]) if want_idx else None, _lib.stream()))

    if n_loc > 0 and (on_device or n_loc * D * 8 < (32 << 20)):
        run(data[r_lo:r_hi] if on_device else _lib.to_dev(data[r_lo:r_hi]), b_lo, b_hi, 0)
    elif n_loc > 0:
        # host input: groups of whole batches (~16 MB) go through two pinned staging
        # buffers, so the upload of the next group overlaps the kernel of this one
        main, side = t.cuda.current_stream(), _lib.side_stream()
        groups, g0 = [], b_lo
        while g0 < b_hi:
            g1 = g0 + 1
            while g1 < b_hi and (batch_off[g1 + 1] - batch_off[g0]) * D * 8 <= (16 << 20):
                g1 += 1
            groups.append((g0, g1))
            g0 = g1
        cap = max(int(batch_off[b] - batch_off[a]) for a, b in groups)
        pins = [_lib.pinned(("knn_in", i), cap * D, t.float64) for i in range(2)]
        busy = [None, None]
        for i, (ga, gb) in enumerate(groups):
            s0, s1 = int(batch_off[ga]), int(batch_off[gb])
            if busy[i % 2] is not None:
                busy[i % 2][0].synchronize()  # the copy out of this staging buffer is done
            stage = pins[i % 2][:(s1 - s0) * D].view(s1 - s0, D)
            stage.numpy()[...] = data[s0:s1]
            with t.cuda.stream(side):
                d_rows = stage.to("cuda", non_blocking=True)
                up = t.cuda.Event()
                up.record(side)
            d_rows.record_stream(main)
            main.wait_event(up)
            run(d_rows, ga, gb, s0 - r_lo)
            busy[i % 2] = (up, d_rows)
    # the ranks' disjoint slices are all-gathered on the device (batches are contiguous,
    # so rank order is row order), then downloaded once
    kth = _dist.allgather_concat_tensor(t_kth)
    if not on_device:  # device input: the k-th distances stay on the device too
        kth = kth.cpu().numpy()
    dist = idx = None
    if want_dist:
        dist = _dist.allgather_concat_tensor(t_dist.view(n_loc, kk)).cpu().numpy()
    if want_idx:
        idx = _dist.allgather_concat_tensor(t_idx.view(n_loc, kk)).cpu().numpy()
    return kth, dist, idx


def bw_knn(data, weights=None, K_eff=100, k=None, batch_size=10000, precision="float64"):
    """K Nearest Neighbors method (reference kde.py:41-103).

    For each sample, the bandwidth is the distance to its k-th neighbour inside
    its batch (``np.array_split`` into ``round(N / batch_size)`` batches), times
    ``f_k ** (1/dim)``.  ``precision="float64"`` reproduces the reference's
    distances bit for bit; ``"float32"`` is the fast mode.
    """
    if not _lib_is_tensor(data):
        data = np.asarray(data)
    N, dim = data.shape
    batches = int(max(1, np.round(N / batch_size)))
    # The reference's branches are inverted (kde.py:70-74): with weights
    # N_eff = N, without weights it raises on np.sum(None ** 2).
    if weights is None:
        raise TypeError("unsupported operand type(s) for ** or pow(): 'NoneType' and 'int'")
    N_eff = N
    if k is None:  # Compute k based on K_eff
        k_float = batch_size * K_eff / N_eff
        k = int(max(1, round(k_float)))
        f_k = k_float / k  # Correction factor for k
    else:  # k set by user
        f_k = 1.0
        K_eff = k * N_eff / batch_size
    print("Using k = {} neighbors per batch (batch_size = {})".format(k, batch_size))
    print("Correction factor: f_k = k_float / k = {}".format(f_k))
    print("Effective total neighbors: K_eff = {}".format(K_eff))
    off = array_split_bounds(N, batches)
    smallest = int(np.min(np.diff(off)))
    if k + 1 > smallest:  # scikit-learn's check in kneighbors (kde.py:96)
        raise ValueError("Expected n_neighbors <= n_samples_fit, but n_neighbors = {}, "
                         "n_samples_fit = {}".format(k + 1, smallest))
    kth, _, _ = knn_batched(data, k + 1, off, precision=precision)
    return kth * (f_k) ** (1 / dim)


# ---------------------------------------------------------------------------
# MLCV
# ---------------------------------------------------------------------------

def kfold_bounds(N, K):
    """Fold boundaries of sklearn ``KFold(n_splits=K)`` without shuffle
    (kde.py:134-136): contiguous, the first ``N % K`` folds one longer."""
    sizes = np.full(K, N // K, dtype=np.int64)
    sizes[: N % K] += 1
    return np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)


class MLCVPlan:
    """Device-resident state of one MLCV grid search: packed sources, the test
    rows of this rank with their fold intervals, and the output buffers.

    ``launch()`` enqueues the kernels for the whole grid on the current stream (no host
    work, no synchronisation): the fp32 pair kernel, the fp64 recomputation of the rows
    it flagged (sums inside the flush-to-zero window of ``ex2.approx``), and the
    fixed-order reduction to one partial log-likelihood per grid point.  ``collect()``
    all-reduces those G doubles over the ranks on the device (NCCL) when asked to, and
    downloads them.  ``dtype="float64"`` evaluates every row with the fp64 kernel
    (1e-12 against the reference's fp64 arithmetic, kde.py:106-153).

    With torch.distributed initialised and ``rows=None`` every rank uploads only its own
    share of the list; the shares are all-gathered on the device (NVLink).
    """

    def __init__(self, data, weights, seed, grid, n_splits=10, rows=None, nsplit=None,
                 dtype="float32"):
        L = _lib.load()
        t = _lib.torch()
        if dtype not in ("float32", "float64"):
            raise ValueError("dtype must be 'float32' or 'float64'")
        self.dtype = dtype
        data = np.ascontiguousarray(data, dtype=np.float64)
        N, D = data.shape
        self.N, self.D = N, D
        self.grid = np.ascontiguousarray(grid, dtype=np.float64).reshape(-1)
        K = self.K = int(min(n_splits, N))
        if K < 2:
            raise ValueError("n_splits must be at least 2")
        if np.ndim(seed) == 1:
            if len(seed) < N:
                raise ValueError("If bw is array, must have same len as data")
            seed_arr = np.ascontiguousarray(seed, dtype=np.float64)[:N]
            seed_scalar = 0.0
            hmin = float(seed_arr.min())
        else:
            seed_arr = None
            seed_scalar = float(seed)
            hmin = seed_scalar
        if not hmin > 0:
            raise ValueError("bandwidths must be positive")
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)[:N]
        fb = self._fb = kfold_bounds(N, K)
        nf = np.diff(fb).astype(np.float64)
        wsum_all = float(np.sum(w)) if w is not None else float(N)
        wfold = np.add.reduceat(w, fb[:-1]) if w is not None else nf
        wtrain = wsum_all - wfold
        cwf = 1.0 / (nf * K)
        self._sum_cw_all = float(np.sum(wfold * cwf))
        rank, world = _dist.rank_world()
        gather = rows is None and world > 1
        if rows is None:
            rows = _dist.shard_range(N, rank, world)
        r_lo, r_hi = self.rows = (int(rows[0]), int(rows[1]))
        M = self.M = r_hi - r_lo
        wmax = float(np.max(w)) if w is not None else 1.0
        self.w_scale = 1.0 / wmax
        self.lc_ref = -D * np.log2(hmin)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.sum_cw = 0.0
        st = _lib.stream()

        def up(a):
            """Upload a per-sample array: only this rank's rows when the ranks share the
            work, then all-gathered on the device."""
            if a is None:
                return None
            if gather:
                self.h2d_bytes += a[r_lo:r_hi].nbytes
                return _dist.allgather_shards(_lib.to_dev(a[r_lo:r_hi]), N)
            self.h2d_bytes += a.nbytes
            return _lib.to_dev(a)

        d_data = self.d_data = up(data)
        d_w = self.d_w = up(w)
        d_bw = self.d_bw = up(seed_arr)
        if M == 0:
            self.chunks = []
            self.n_launches = 0
            return
        # centre of the packed fp32 coordinates: column means, computed on the device
        mom_scr = _lib.empty(L.kdsb200_moments_scratch(), t.float64)
        mom = _lib.empty(D + 2, t.float64)
        _lib.check(L.kdsb200_weighted_moments(_lib.ptr(d_data), None, N, D, None, 1,
                                              _lib.ptr(mom_scr), _lib.ptr(mom), st))
        # fold bookkeeping of the rows, on the device from the K-entry fold tables
        Mpad = self.Mpad = L.kdsb200_mlcv_mpad(M)
        with np.errstate(divide="ignore"):
            lwt = np.log(wtrain)
        d_cwf, d_lwt = _lib.to_dev(cwf), _lib.to_dev(lwt)
        self.h2d_bytes += cwf.nbytes + lwt.nbytes
        self.d_lo, self.d_hi = _lib.empty(Mpad, t.int32), _lib.empty(Mpad, t.int32)
        self.d_cw, self.d_lw = _lib.empty(Mpad, t.float64), _lib.empty(Mpad, t.float64)
        _lib.check(L.kdsb200_mlcv_rows_setup(N, K, r_lo, M, Mpad, _lib.ptr(d_w), _lib.ptr(d_cwf),
                                             _lib.ptr(d_lwt), _lib.ptr(self.d_lo),
                                             _lib.ptr(self.d_hi), _lib.ptr(self.d_cw),
                                             _lib.ptr(self.d_lw), st))
        mom_cw = _lib.empty(3, t.float64)
        _lib.check(L.kdsb200_weighted_moments(_lib.ptr(self.d_cw), None, M, 1, None, 1,
                                              _lib.ptr(mom_scr), _lib.ptr(mom_cw), st))
        center = mom.cpu().numpy()[:D] / N
        self.sum_cw = float(mom_cw.cpu()[0])
        cen_keep, cen_p = _lib.hostvec(center)
        # fp64 coefficients: used by the fp64 mode and by the recomputation of flagged rows
        self.d_coef = _lib.empty(2 * N, t.float64)
        _lib.check(L.kdsb200_mlcv_coef_f64(_lib.ptr(d_w), _lib.ptr(d_bw), seed_scalar, N, D,
                                           self.w_scale, _lib.ptr(self.d_coef), st))
        self.log_shift = -self.lc_ref * np.log(2.0)  # fp32 sums are in units of 2^lc_ref
        if dtype == "float32":
            self.packed = _lib.empty(L.kdsb200_packed_floats(N, D), t.float32)
            _lib.check(L.kdsb200_pack_sources_f32(
                _lib.ptr(d_data), _lib.ptr(d_w), _lib.ptr(d_bw), seed_scalar, N, D, cen_p,
                self.w_scale, self.lc_ref, 1, _lib.ptr(self.packed), st))
            self.qT = _lib.empty(Mpad * D, t.float32)
            d_rows = d_data[r_lo:r_hi]  # contiguous row slice of the (N,D) tensor
            _lib.check(L.kdsb200_pack_queries_f32(_lib.ptr(d_rows), M, Mpad, D, cen_p,
                                                  _lib.ptr(self.qT), st))
            # source-dimension split (None: the library picks the one that fills whole waves)
            self.nsplit = int(nsplit) if nsplit else L.kdsb200_mlcv_nsplit(N, M)
            self.nsplit = max(1, min(self.nsplit, (N + 511) // 512))
            self.flagmask = _lib.empty(Mpad, t.int32)
        else:
            self.nsplit = 1
            self.flagmask = None
        self.nblk = Mpad // 256
        self.nb64 = int(L.kdsb200_mlcv_f64_blocks(M))
        self.d_scores = _lib.empty(len(self.grid), t.float64)
        self.red_tmp = _lib.empty(L.kdsb200_mlcv_reduce_scratch(), t.float64)
        self.fix_scratch = (_lib.empty(L.kdsb200_mlcv_fix_scratch_bytes(M), t.uint8)
                            if dtype == "float32" else None)
        self.chunks = []
        for g0 in range(0, len(self.grid), 32):
            gsub = np.ascontiguousarray(self.grid[g0:g0 + 32])
            gt = L.kdsb200_mlcv_gtile(len(gsub))
            gt64 = L.kdsb200_mlcv_f64_gtile(len(gsub))
            nk = np.ascontiguousarray(-_LOG2E / (2.0 * gsub ** 2), dtype=np.float32)
            f32 = dtype == "float32"
            sbuf = _lib.empty(self.nsplit * Mpad * gt, t.float64) if self.nsplit > 1 else None
            partial = _lib.empty(self.nblk * gt, t.float64) if f32 else None
            partial64 = _lib.empty(self.nb64 * gt64, t.float64)
            self.chunks.append((g0, gsub, gt, gt64, nk, sbuf, partial, partial64))
        #: kernel launches per launch() call
        # pair kernel (+ finalize), slot assignment + two fp64 row launches + combine, or the
        # fp64 row kernel alone; then the two-stage reduction
        self.n_launches = len(self.chunks) * ((7 if self.nsplit == 1 else 8)
                                              if dtype == "float32" else 3)

    def _row_folds(self):
        nf = np.diff(self._fb).astype(np.float64)
        fold_of = np.repeat(np.arange(self.K), np.diff(self._fb))
        return nf, fold_of[self.rows[0]:self.rows[1]]

    @property
    def n_evals(self):
        """Kernel evaluations (exponentials that enter a score) per launch() call."""
        if self.M == 0:
            return 0.0
        nf, f_loc = self._row_folds()
        return float(np.sum(float(self.N) - nf[f_loc])) * len(self.grid)

    @property
    def n_slots(self):
        """(row, source, g) slots the fp32 kernel executes per launch(): per CTA of 256
        rows, all source tiles except those inside the excluded interval shared by all
        its rows (benchmark accounting; not needed to run the search)."""
        if self.M == 0:
            return 0.0
        _, f_loc = self._row_folds()
        fb = self._fb
        N, M, Mpad = self.N, self.M, self.Mpad
        TS = 512
        ntiles = (N + TS - 1) // TS
        lo_pad = np.full(Mpad, -1, dtype=np.int64)
        hi_pad = np.full(Mpad, np.iinfo(np.int64).max, dtype=np.int64)
        lo_pad[:M] = fb[f_loc]
        hi_pad[:M] = fb[f_loc + 1]
        cta_lo = lo_pad.reshape(-1, 256).max(axis=1)
        cta_hi = hi_pad.reshape(-1, 256).min(axis=1)
        skipped = np.clip(cta_hi // TS - (cta_lo + TS - 1) // TS, 0, None)
        skipped[cta_hi <= cta_lo] = 0
        return float(np.sum(ntiles - skipped)) * TS * 256 * len(self.grid)

    def launch(self):
        if self.M == 0:
            return
        L = _lib.load()
        st = _lib.stream()
        r_lo = self.rows[0]
        for g0, gsub, gt, gt64, nk, sbuf, partial, partial64 in self.chunks:
            G = len(gsub)
            if self.dtype == "float32":
                _lib.check(L.kdsb200_mlcv_scores_f32(
                    _lib.ptr(self.packed), self.N, self.D, _lib.ptr(self.qT), self.M, self.Mpad,
                    _lib.ptr(self.d_lo), _lib.ptr(self.d_hi), _lib.ptr(self.d_cw),
                    _lib.ptr(self.d_lw), nk.ctypes.data_as(_lib.c_float_p), G,
                    self.nsplit, _lib.ptr(sbuf), _lib.ptr(partial), _lib.ptr(self.flagmask), st))
            _lib.check(L.kdsb200_mlcv_rows_f64(
                _lib.ptr(self.d_data), _lib.ptr(self.d_coef), self.N, self.D, r_lo, self.M,
                _lib.ptr(self.d_lo), _lib.ptr(self.d_hi), _lib.ptr(self.d_cw), _lib.ptr(self.d_lw),
                gsub.ctypes.data_as(_lib.c_double_p), G, self.log_shift,
                _lib.ptr(self.flagmask), _lib.ptr(self.fix_scratch), _lib.ptr(partial64), st))
            out = self.d_scores[g0:g0 + G]
            if self.dtype == "float32":
                _lib.check(L.kdsb200_mlcv_reduce(_lib.ptr(partial), self.nblk, gt,
                                                 _lib.ptr(partial64), self.nb64, gt64, G,
                                                 _lib.ptr(self.red_tmp), _lib.ptr(out), st))
            else:
                _lib.check(L.kdsb200_mlcv_reduce(_lib.ptr(partial64), self.nb64, gt64, None, 0, 0,
                                                 G, _lib.ptr(self.red_tmp), _lib.ptr(out), st))

    def n_flagged(self):
        """Rows of the last launch() whose fp32 sums were recomputed in fp64 (synchronises)."""
        if self.M == 0 or self.flagmask is None:
            return 0
        return int((self.flagmask[:self.M] != 0).sum().item())

    def collect(self, allreduce=False):
        """Scores contributed by the rows of this plan; ``allreduce=True`` sums the G
        partial log-likelihoods over the ranks on the device first (the ranks' rows must
        partition the list, as they do with ``rows=None``)."""
        G = len(self.grid)
        t = _lib.torch()
        D = self.D
        const_g = (self.lc_ref * np.log(2.0) - np.log(self.w_scale)
                   - 0.5 * D * np.log(2 * np.pi) - D * np.log(self.grid))
        if allreduce and _dist.rank_world()[1] > 1:
            if self.M == 0:
                self.d_scores = t.zeros(G, dtype=t.float64, device="cuda")
            _dist.allreduce_sum_tensor(self.d_scores)
            part = self.d_scores.cpu().numpy()
            self.d2h_bytes = G * 8
            with np.errstate(invalid="ignore"):
                return part + const_g * self._sum_cw_all
        if self.M == 0:
            return np.zeros(G)
        part = self.d_scores.cpu().numpy()
        self.d2h_bytes = G * 8
        with np.errstate(invalid="ignore"):
            return part + const_g * self.sum_cw


def mlcv_scores(data, weights, seed, grid, n_splits=10, dtype="float32"):
    """Cross-validation scores of ``bw_grid[g] = grid[g] * seed`` for every g.

    ``score[g] == _kde_cv_score(grid[g] * seed, data, weights, n_splits)`` of
    the reference (kde.py:106-153), computed for the whole grid in one pass
    over the pair space.  ``seed`` is a scalar or one bandwidth per sample.
    Test rows are sharded over ranks when torch.distributed is initialised.
    ``dtype="float32"`` is within 1e-5 of the reference's fp64 arithmetic (rows whose
    sums fall into the fp32 underflow window are recomputed in fp64), ``"float64"``
    within 1e-12.
    """
    plan = MLCVPlan(data, weights, seed, grid, n_splits=n_splits, dtype=dtype)
    plan.launch()
    return plan.collect(allreduce=True)


def _kde_cv_score(bw, data, weights=None, n_splits=10, modelKDE=TreeKDE, dtype="float32"):
    """Cross-validation likelihood score (reference kde.py:106-153).

    ``modelKDE`` is accepted for signature compatibility; the score is always
    evaluated with the Gaussian kernel, as in the reference (kde.py:146).
    """
    if np.ndim(bw) == 1:
        if len(bw) < len(data):
            raise ValueError("If bw is array, must have same len as data")
        if len(bw) > len(data):
            print("Warning: bw longer than data.")
    return float(mlcv_scores(data, weights, bw, np.array([1.0]), n_splits=n_splits,
                             dtype=dtype)[0])


def _effective_n(n_rows, weights):
    if weights is None:
        return n_rows
    weights = np.asarray(weights)
    return np.sum(weights) ** 2 / np.sum(weights ** 2)


def _maybe_plot_scores(grid, scores):
    """Score-vs-factor plot of the reference (kde.py:214-220); skipped when
    matplotlib is not installed."""
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        return
    plt.plot(grid, scores)
    plt.xlabel("Scaling factor")
    plt.ylabel("MLCV Figure of Merit (FoM)")
    plt.tight_layout()
    plt.show()


def bw_mlcv(data, weights=None, n_splits=10, seed=None, grid=None, show=True, dtype="float32"):
    """Maximum Likelihood Cross-Validation bandwidth (reference kde.py:156-226).

    ``bw_grid[g] = grid[g] * seed`` (seed scalar or per-sample, Silverman by
    default; default grid ``logspace(-0.1, 0.1, 10)``); returns the grid entry
    with the best K-fold score and raises if the best one is the first or last.
    """
    data = np.asarray(data)
    if seed is None:
        seed = bw_silv(data.shape[1], _effective_n(len(data), weights))
    if grid is None:
        print("No grid specified. Using np.logspace(-0.1, 0.1, 10).")
        print("If fitting fails, change the grid.")
        grid = np.logspace(-0.1, 0.1, 10)
    factors = np.asarray(grid, dtype=np.float64).reshape(-1)
    scores = mlcv_scores(data, weights, seed, factors, n_splits=min(n_splits, len(data)),
                         dtype=dtype)
    best = int(np.argmax(scores))
    if show:
        _maybe_plot_scores(factors, scores)
    if best == 0 or best == len(factors) - 1:
        raise Exception(
            "Maximum not found in bw range selected. Move grid and try again."
        )
    return factors[best] * seed


def bw_auto(data, weights, grid=None, Nmlcv=1e4, Nbatch=1e4, Nknn=10, show=True,
            n_splits=10):
    """kNN seed + MLCV rescale (reference kde.py:229-297).

    1. adaptive ``bw_knn`` on all samples (``Nbatch=-1``: one batch);
    2. ``bw_mlcv`` on the first ``Nmlcv`` samples seeded with (1) (``-1``: all);
    3. the MLCV factor is applied to the whole kNN bandwidth and corrected by
       the Silverman ratio between N and Nmlcv samples.
    """
    on_device = _lib_is_tensor(data)
    if not on_device:
        data = np.asarray(data)
    n_all, dim = data.shape
    print("Fitting first with kNN method:")
    if Nbatch == -1:
        print("No batch size for kNN selected. Using all the particles list.")
        Nbatch = n_all
    knn_bw = bw_knn(data, weights=weights, k=Nknn, batch_size=Nbatch)
    print("")
    print("Fitting now with MLCV using the previous fitting as seed:")
    if Nmlcv == -1:
        print("No size for MLCV selected. Using all the particles list.")
        Nmlcv = n_all
    n_cv = int(Nmlcv)
    print("If fitting takes too long, consider reducing Nmlcv.")
    if on_device:
        # the cross-validation subset (Nmlcv rows) goes through the host-array entry point;
        # the full-length kNN bandwidths stay on the device
        w_cv = weights[:n_cv].cpu().numpy() if _lib_is_tensor(weights) else weights
        cv_bw = bw_mlcv(data[:n_cv].cpu().numpy(), weights=w_cv,
                        seed=knn_bw[:n_cv].cpu().numpy(), grid=grid, show=show,
                        n_splits=n_splits)
        first = float(knn_bw[0].item())
    else:
        cv_bw = bw_mlcv(data[:n_cv], weights=weights, seed=knn_bw[:n_cv], grid=grid,
                        show=show, n_splits=n_splits)
        first = knn_bw[0]
    print("")
    print("Extending the MLCV optimization to full kNN bandwidth:")
    factor = cv_bw[0] / first
    factor *= bw_silv(dim, len(knn_bw)) / bw_silv(dim, len(cv_bw))
    print("Done.")
    return knn_bw * factor


def optimize_bw(bw_method, vecs, ws=None, weightfun=None, maskfun=None, **kwargs):
    """Dispatch to a bandwidth method after the reference's sample filtering
    (kde.py:300-354): drop non-positive weights, apply ``weightfun`` to the
    weights and ``maskfun`` to the samples."""
    if bw_method not in bw_methods:
        raise Exception("Invalid bw_method. Available: {}".format(list(bw_methods)))
    vecs = np.array(vecs)
    keep = np.ones(len(vecs), dtype=bool)
    if ws is not None:
        ws = np.array(ws)
        keep = ws > 0
    if weightfun is not None:
        ws = ws * weightfun(vecs)
    if maskfun is not None:
        keep = keep & maskfun(vecs)
    vecs = vecs[keep]
    ws = ws[keep]  # like the reference, ws=None is not supported here
    if bw_method == "silv":
        return bw_silv(vecs.shape[1], _effective_n(len(vecs), ws))
    return bw_methods[bw_method](vecs, weights=ws, **kwargs)


bw_methods = {"auto": bw_auto, "silv": bw_silv, "knn": bw_knn, "mlcv": bw_mlcv}
