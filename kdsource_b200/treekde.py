"""GPU kernel-density estimator with the interface KDSource uses from KDEpy.

The reference builds ``KDEpy.TreeKDE(kernel, bw)`` and calls ``.fit(data,
weights=ws)`` / ``.evaluate(points)`` (python/kdsource/kdsource.py:125,
211-212, 239; kde.py:146-151) and reads/assigns ``.bw``, ``.data``,
``.weights`` (kdsource.py:244, 288, 457-479).  :class:`TreeKDE` here offers
exactly that surface, backed by the ``kde_eval`` CUDA kernels through the C
ABI (``kdsb200_kde_eval_f32`` / ``_f64``).

Semantics: ``out[m] = sum_i w_i/sum(w) (2 pi)^(-d/2) h_i^-d exp(-|q_m - x_i|^2 /
(2 h_i^2))``.  ``cutoff=None`` is the exact sum (KDEpy NaiveKDE semantics);
``cutoff="treekde"`` applies the hard support radius KDEpy's TreeKDE derives
from the largest bandwidth; a float is an explicit radius; ``cutoff="adaptive"``
cuts every source at the radius that rule gives relative to its own bandwidth
(r_i = c h_i), which is what keeps the truncation local for per-sample bandwidths.

``dtype``: ``"float32"`` -- sources ordered along a Hilbert curve and stored in tiles of
512 with coordinates relative to the tile centre, so the fp32 rounding error scales with
the tile size instead of the spread of the list; with a cutoff, tiles farther than the
radius from a block of 1024 (equally ordered) query points are skipped, exactly;
``"float32x2"`` (Gaussian kernel; coordinates carried as hi + lo floats, error
independent of spread / bandwidth, ~57 % of the fp32 rate); ``"float32_dense"`` (the
unordered fp32 kernel: 1e-5 only while the smallest bandwidth is above ~0.28 of the
spread); ``"float64"`` (1e-12); ``"auto"`` (default): float32, switching to float32x2
when the smallest bandwidth is below 0.07 of the largest tile half-extent.

``kernel="epa"`` / ``"box"`` (the other two names the reference passes to KDEpy,
kdsource.py:60-65) are KDEpy's radially symmetric finite-support kernels scaled
to unit variance: support radius ``R_i = h_i sqrt(5)`` (epa) or ``h_i sqrt(3)``
(box), ``K = (d+2)/(2 V_d R^d) (1 - r^2/R^2)`` resp. ``1/(V_d R^d)`` inside the
ball (``V_d`` = volume of the unit d-ball).
"""

import ctypes

import numpy as np

from . import _lib

_LOG2E = 1.4426950408889634
_KERNEL_ID = {"gaussian": 0, "epa": 1, "box": 2}
_SUPPORT_SCALE = {"epa": np.sqrt(5.0), "box": np.sqrt(3.0)}  # 1/sqrt(var), KDEpy kernel table


def unit_ball_volume(d):
    from math import gamma, pi

    return pi ** (d / 2.0) / gamma(d / 2.0 + 1.0)


_KERNEL_ALIASES = {"g": "gaussian", "gaussian": "gaussian", "e": "epa", "epa": "epa",
                   "b": "box", "box": "box"}


def treekde_radius(bw_max, atol=1e-3):
    """Support radius of KDEpy's TreeKDE for a Gaussian kernel: the root of the
    1-D pdf N(0, bw_max) = atol, bracketed on [0, 8 bw_max], plus xtol=1e-3."""
    c = bw_max * np.sqrt(2 * np.pi) * atol
    if c >= 1.0:
        return 1e-3
    return float(bw_max * np.sqrt(-2.0 * np.log(c)) + 1e-3)


class TreeKDE:
    #: dtype="auto" switches to the split-coordinate kernel when the smallest bandwidth is
    #: below this fraction of the largest half-extent of a source tile (the fp32 error of a
    #: kernel value is ~ (extent / h) (r / h) 1.2e-7; 0.07 keeps it under 1e-5 out to r = 6 h)
    AUTO_SPLIT_BELOW = 0.07

    def __init__(self, kernel="gaussian", bw=1.0, cutoff=None, dtype="auto"):
        if kernel not in _KERNEL_ALIASES:
            raise ValueError("Unknown kernel {!r}".format(kernel))
        self.kernel = _KERNEL_ALIASES[kernel]
        if isinstance(bw, np.ndarray) and np.ndim(bw) >= 2:
            raise ValueError("bw must be a scalar or a 1-D array")
        self.bw = bw
        self.cutoff = cutoff
        if dtype not in ("auto", "float32", "float32x2", "float32_dense", "float64"):
            raise ValueError("dtype must be 'auto', 'float32', 'float32x2', 'float32_dense' "
                             "or 'float64'")
        self.dtype = dtype
        self._data = None      # host copies (numpy) ...
        self._weights = None
        self._d_data = None    # ... and / or device tensors (fit_device)
        self._d_w = None
        self._d_bw = None      # device copy of a per-sample bandwidth array, with the
        self._d_bw_for = None  # host array it mirrors
        self._has_w = False
        self._gen = 0          # bumped whenever the samples change
        self._dev = None

    # `.data` / `.weights` are what the reference reads from KDEpy (kdsource.py:457-479);
    # after fit_device they are downloaded on first access
    @property
    def data(self):
        if self._data is None and self._d_data is not None:
            self._data = self._d_data.cpu().numpy()
        return self._data

    @data.setter
    def data(self, value):
        self._data = value
        self._d_data = None
        self._gen += 1

    @property
    def weights(self):
        if self._weights is None and self._d_w is not None:
            self._weights = self._d_w.cpu().numpy()
        return self._weights

    @weights.setter
    def weights(self, value):
        self._weights = value
        self._d_w = None
        self._has_w = value is not None
        self._gen += 1

    def fit_device(self, d_data, d_weights=None, d_bw=None):
        """:meth:`fit` from (N,D) / (N,) float64 CUDA tensors that stay on the device
        (the device-resident fit of KDSource); `d_bw` is an optional device copy of a
        per-sample ``bw`` array."""
        t = _lib.torch()
        if d_data.dim() != 2 or d_data.shape[0] == 0 or d_data.shape[1] > 8:
            raise ValueError("data must have shape (obs, dims) with obs > 0 and dims <= 8")
        self._data = None
        self._weights = None
        self._d_data = d_data.to(dtype=t.float64).contiguous()
        self._d_w = None if d_weights is None else d_weights.to(dtype=t.float64).contiguous()
        self._has_w = d_weights is not None
        if d_bw is not None:
            self._d_bw, self._d_bw_for = d_bw, self.bw
        self._gen += 1
        self._dev = None
        return self

    def n_samples(self):
        if self._d_data is not None:
            return int(self._d_data.shape[0]), int(self._d_data.shape[1])
        return self._data.shape

    # -- KDEpy interface ----------------------------------------------------
    def fit(self, data, weights=None):
        data = np.ascontiguousarray(data, dtype=np.float64)
        if data.ndim == 1:
            data = data.reshape(-1, 1)
        if data.ndim != 2 or len(data) == 0:
            raise ValueError("data must have shape (obs, dims) with obs > 0")
        if data.shape[1] > 8:
            raise ValueError("at most 8 dimensions are supported")
        if weights is not None:
            weights = np.ascontiguousarray(weights, dtype=np.float64).reshape(-1)
            if len(weights) != len(data):
                raise ValueError("weights must have the same length as data")
            if np.any(weights < 0) or not np.sum(weights) > 0:
                raise ValueError("weights must be non-negative with positive sum")
        self.data = data
        self.weights = weights
        self._dev = None
        return self

    def __call__(self, points):
        return self.evaluate(points)

    def evaluate(self, points, chunk=None, shard=True):
        """Densities at `points` ((obs, dims)).  With torch.distributed initialised
        and ``shard=True`` the query points are split contiguously over the ranks
        (sources replicated) and the result is assembled on every rank.

        The points travel in pieces through two pinned staging buffers: the upload of piece
        i+1 and the download of piece i-1 overlap the kernels of piece i."""
        if self._data is None and self._d_data is None:
            raise ValueError("Must call fit before evaluating.")
        D = self.n_samples()[1]
        pts = np.ascontiguousarray(points, dtype=np.float64)
        if pts.ndim == 1:
            pts = pts.reshape(-1, D)
        if pts.shape[1] != D:
            raise ValueError("points must have shape (obs, {})".format(D))
        from . import dist as _dist

        rank, world = _dist.rank_world()
        lo, hi = _dist.shard_range(len(pts), rank, world) if shard else (0, len(pts))
        n = hi - lo
        dev = self._device_state()
        if chunk is None:  # at least four pieces once the input is large enough to overlap
            chunk = int(min(1 << 22, max(1 << 16, -(-n // 4))))
        culled = dev["dtype"] == "float32"

        def work(d_pts, m):
            d_out = self.evaluate_device(d_pts, m, dev)
            return (d_out, self.last_tiles_visited) if culled else (d_out,)

        (out,), visited = _lib.stream_pieces(pts, lo, hi, chunk, work, tag="eval")
        if culled and n:  # tiles loaded over all pieces of this call
            t = _lib.torch()
            self.last_tiles_visited = t.tensor(sum(float(v.item()) for v in visited),
                                               dtype=t.float64)
        if shard and world > 1:  # the ranks' disjoint slices, in rank (= query) order
            out = _dist.allgather_rows(out)
        return out

    # -- device side ----------------------------------------------------------
    def _bw_array(self):
        N = self.n_samples()[0]
        if np.ndim(self.bw) == 0:
            return None, float(self.bw)
        bw = np.ascontiguousarray(self.bw, dtype=np.float64).reshape(-1)
        if len(bw) < N:
            raise ValueError("If bw is array, must have same len as data")
        return bw[:N], 0.0

    # Has the per-sample bandwidth array changed since the sources were packed (users and
    # KDSource.save rescale it in place)?  Small arrays are compared in full; beyond 2^20
    # entries a strided sample of 4096 values stands in for the array.
    @staticmethod
    def _bw_snapshot(bw):
        if len(bw) <= (1 << 20):
            return bw.copy()
        return (len(bw), bw[::max(1, len(bw) // 4096)].copy())

    @staticmethod
    def _bw_same(bw, snap):
        if isinstance(snap, tuple):
            return len(bw) == snap[0] and np.array_equal(bw[::max(1, len(bw) // 4096)], snap[1])
        return len(bw) == len(snap) and np.array_equal(bw, snap)

    def _cutoff_radius(self, bw_max):
        if self.cutoff is None or self.cutoff == "adaptive":
            return 0.0
        if isinstance(self.cutoff, str):
            if self.cutoff != "treekde":
                raise ValueError("cutoff must be None, 'treekde', 'adaptive' or a radius")
            return treekde_radius(bw_max)
        return float(self.cutoff)

    def _cutoff_rel(self, bw_min):
        """cutoff="adaptive": every source is cut at r_i = c h_i, with c the ratio KDEpy's
        rule gives for the smallest bandwidth (so no source is cut closer than TreeKDE would
        cut it).  The same as "treekde" for one bandwidth; with per-sample bandwidths the
        radius follows the local bandwidth instead of the largest one of the whole list."""
        if self.cutoff != "adaptive":
            return 0.0
        return treekde_radius(bw_min) / bw_min

    def _device_state(self):
        """Upload and pack the fitted sources (cached until data/bw change)."""
        bw_arr, bw_scalar = self._bw_array()
        d = self._dev
        if (d is not None and d["gen"] == self._gen
                and d["dtype_spec"] == self.dtype and d["cutoff_spec"] == self.cutoff
                and d["kernel"] == self.kernel
                and d["bw_scalar_snap"] == bw_scalar
                and ((bw_arr is None) == (d["bw_snap"] is None))
                and (bw_arr is None or self._bw_same(bw_arr, d["bw_snap"]))):
            return d
        L = _lib.load()
        t = _lib.torch()
        N, D = self.n_samples()
        hmin = float(np.min(bw_arr)) if bw_arr is not None else bw_scalar
        hmax = float(np.max(bw_arr)) if bw_arr is not None else bw_scalar
        if not hmin > 0:
            raise ValueError("bandwidths must be positive")
        if self._d_data is not None:
            # device-resident samples: statistics from the reduction kernels
            d_data, d_w = self._d_data, self._d_w
            if d_w is not None:
                sums, sumw, _ = _lib.dev_moments(d_data, d_w, N, D)
                wmax = float(_lib.dev_minmax(d_w, N, 1)[1][0])
            else:
                sums, sumw, _ = _lib.dev_moments(d_data, None, N, D)
                wmax = 1.0
            center = sums / sumw
            if (bw_arr is not None and self._d_bw is not None and self._d_bw_for is self.bw
                    and len(self._d_bw) >= N):
                d_bw = self._d_bw[:N]
            else:
                d_bw = _lib.to_dev(bw_arr) if bw_arr is not None else None
        else:
            w = self._weights
            sumw = float(np.sum(w)) if w is not None else float(N)
            wmax = float(np.max(w)) if w is not None else 1.0
            if w is not None:
                center = np.average(self._data, axis=0, weights=w)
            else:
                center = self._data.mean(axis=0)
            d_data = _lib.to_dev(self._data)
            d_w = _lib.to_dev(w) if w is not None else None
            d_bw = _lib.to_dev(bw_arr) if bw_arr is not None else None
        gauss = self.kernel == "gaussian"
        dtype = self.dtype
        dev = {"N": N, "D": D, "center": center, "kernel": self.kernel, "dtype_spec": self.dtype,
               "cutoff": self._cutoff_radius(hmax) if gauss else 0.0, "gen": self._gen,
               "cutoff_spec": self.cutoff, "bw_scalar_snap": bw_scalar,
               "bw_snap": None if bw_arr is None else self._bw_snapshot(bw_arr)}
        w_scale = 1.0 / sumw
        st = _lib.stream()
        if dtype in ("auto", "float32"):
            # ordered tiles: column ranges -> Hilbert order -> tiles with boxes and centres
            mm_scr = _lib.empty(L.kdsb200_minmax_scratch(), t.float64)
            mm = _lib.empty(2 * D, t.float64)
            _lib.check(L.kdsb200_column_minmax(_lib.ptr(d_data), N, D, _lib.ptr(mm_scr),
                                               _lib.ptr(mm), st))
            mmh = mm.cpu().numpy()
            lo_keep, lo_p = _lib.hostvec(mmh[:D])
            hi_keep, hi_p = _lib.hostvec(mmh[D:])
            maxabs = float(max(np.max(np.abs(mmh[:D] - center)), np.max(np.abs(mmh[D:] - center)),
                               1e-300))
            # hi parts and tile centres live on the grid 2^-S with |x - c| 2^S < 2^20 (spare
            # bits for query points outside the sources' bounding box)
            grid = float(2.0 ** (20 - int(np.ceil(np.log2(maxabs)))))
            perm = _lib.empty(N, t.int32)
            scr = _lib.empty(L.kdsb200_order_scratch_bytes(N), t.uint8)
            _lib.check(L.kdsb200_hilbert_order(_lib.ptr(d_data), N, D, lo_p, hi_p, _lib.ptr(scr),
                                               _lib.ptr(perm), st))
            del scr
            ntl = (N + 511) // 512
            sc = 1.0 if gauss else _SUPPORT_SCALE[self.kernel]
            d_h = d_bw if gauss or d_bw is None else _lib.to_dev(bw_arr * sc)
            lc_ref = np.log2(wmax * w_scale) - D * np.log2(hmin * sc)
            packed = _lib.empty(L.kdsb200_packed_floats(N, D), t.float32)
            tb_stride = int(L.kdsb200_sorted_tilebox_floats(D))
            tilebox = _lib.empty(ntl * tb_stride, t.float32)
            tilecen = _lib.empty(ntl * D, t.float32)
            tilehmax = _lib.empty(ntl, t.float32)
            cen_keep, cen_p = _lib.hostvec(center)
            _lib.check(L.kdsb200_pack_sources_sorted_f32(
                _lib.ptr(d_data), _lib.ptr(d_w), _lib.ptr(d_h), bw_scalar * sc, N, D,
                _lib.ptr(perm), cen_p, grid, w_scale, lc_ref, 0 if gauss else 1,
                _lib.ptr(packed), _lib.ptr(tilebox), _lib.ptr(tilecen), _lib.ptr(tilehmax), st))
            box = tilebox.view(ntl, tb_stride)[:, :2 * D].cpu().numpy().reshape(ntl, 2, D)
            extent = float(np.max(box[:, 1] - box[:, 0])) / 2 if N > 1 else 0.0
            if (dtype == "auto" and gauss and self.cutoff is None
                    and hmin < self.AUTO_SPLIT_BELOW * extent):
                dtype = "float32x2"  # tiles too wide for this bandwidth: split coordinates
            else:
                dtype = "float32"
                if gauss:
                    norm = (2 * np.pi) ** (-D / 2)
                else:
                    norm = (0.5 * (D + 2) if self.kernel == "epa" else 1.0) / unit_ball_volume(D)
                dev.update(packed=packed, tilebox=tilebox, tilecen=tilecen, tilehmax=tilehmax,
                           cull_rel=self._cutoff_rel(hmin) if gauss else 0.0, grid=grid,
                           lo=mmh[:D].copy(), hi=mmh[D:].copy(), tile_extent=extent,
                           out_scale=float(2.0 ** lc_ref * norm),
                           cull=dev["cutoff"] if gauss else hmax * sc)
            del perm
        dev["dtype"] = dtype
        if dtype == "float32x2":
            if not gauss:
                raise ValueError("dtype='float32x2' is implemented for the gaussian kernel")
            if dev["cutoff"]:
                raise ValueError("dtype='float32x2' has no cutoff mode")
            # hi parts live on the grid 2^-S, chosen so that |x - c| 2^S < 2^20 (two spare
            # bits for query points outside the sources' bounding box)
            mlo, mhi = _lib.dev_minmax(d_data, N, D)
            maxabs = float(max(np.max(np.abs(mlo - center)), np.max(np.abs(mhi - center))))
            S = 20 - int(np.ceil(np.log2(max(maxabs, 1e-300))))
            dev["grid"] = float(2.0 ** S)
            lc_ref = np.log2(wmax * w_scale) - D * np.log2(hmin)
            packed = _lib.empty(L.kdsb200_packed_split_floats(N, D), t.float32)
            cen_keep, cen_p = _lib.hostvec(center)
            _lib.check(L.kdsb200_pack_sources_split_f32(
                _lib.ptr(d_data), _lib.ptr(d_w), _lib.ptr(d_bw), bw_scalar, N, D, cen_p,
                w_scale, lc_ref, dev["grid"], _lib.ptr(packed), _lib.stream()))
            dev["packed"] = packed
            dev["out_scale"] = float(2.0 ** lc_ref * (2 * np.pi) ** (-D / 2))
        elif dtype == "float32":
            pass  # packed above
        elif dtype == "float32_dense" and gauss:
            lc_ref = np.log2(wmax * w_scale) - D * np.log2(hmin)
            packed = _lib.empty(L.kdsb200_packed_floats(N, D), t.float32)
            cen_keep, cen_p = _lib.hostvec(center)
            _lib.check(L.kdsb200_pack_sources_f32(
                _lib.ptr(d_data), _lib.ptr(d_w), _lib.ptr(d_bw), bw_scalar, N, D, cen_p,
                w_scale, lc_ref, 0, _lib.ptr(packed), _lib.stream()))
            dev["packed"] = packed
            dev["out_scale"] = float(2.0 ** lc_ref * (2 * np.pi) ** (-D / 2))
        elif dtype == "float32_dense":
            # finite-support kernels: linear coefficients w_i R_i^-D / 2^lc_ref (mode 1)
            # with the support radius R_i in place of the bandwidth
            sc = _SUPPORT_SCALE[self.kernel]
            d_R = _lib.to_dev(bw_arr * sc) if bw_arr is not None else None
            lc_ref = np.log2(wmax * w_scale) - D * np.log2(hmin * sc)
            packed = _lib.empty(L.kdsb200_packed_floats(N, D), t.float32)
            cen_keep, cen_p = _lib.hostvec(center)
            _lib.check(L.kdsb200_pack_sources_f32(
                _lib.ptr(d_data), _lib.ptr(d_w), _lib.ptr(d_R), bw_scalar * sc, N, D, cen_p,
                w_scale, lc_ref, 1, _lib.ptr(packed), _lib.stream()))
            norm = (0.5 * (D + 2) if self.kernel == "epa" else 1.0) / unit_ball_volume(D)
            dev["packed"] = packed
            dev["out_scale"] = float(2.0 ** lc_ref * norm)
        else:
            dev.update(d_data=d_data, d_w=d_w, d_bw=d_bw, bw_scalar=bw_scalar,
                       w_scale=w_scale)
        self._dev = dev
        return dev

    def evaluate_device(self, d_pts, M, dev=None):
        """Densities for M query points already on the device ((M,D) fp64
        tensor); returns an (M,) fp64 device tensor."""
        L = _lib.load()
        t = _lib.torch()
        if dev is None:
            dev = self._device_state()
        N, D = dev["N"], dev["D"]
        d_out = _lib.empty(M, t.float64)
        if M == 0:
            return d_out
        if dev["dtype"] == "float32x2":
            Mpad = L.kdsb200_eval_mpad(M)
            qT = _lib.empty(Mpad * 2 * D, t.float32)
            cen_keep, cen_p = _lib.hostvec(dev["center"])
            _lib.check(L.kdsb200_pack_queries_split_f32(_lib.ptr(d_pts), M, Mpad, D, cen_p,
                                                        dev["grid"], _lib.ptr(qT), _lib.stream()))
            nsplit = L.kdsb200_eval_nsplit(N, M)
            partial = _lib.empty(nsplit * Mpad, t.float64)
            _lib.check(L.kdsb200_kde_eval_split_f32(
                _lib.ptr(dev["packed"]), N, D, _lib.ptr(qT), M, Mpad, dev["out_scale"], nsplit,
                _lib.ptr(partial), _lib.ptr(d_out), _lib.stream()))
        elif dev["dtype"] == "float32":
            st = _lib.stream()
            Mpad = L.kdsb200_sorted_qpad(M)
            lo_keep, lo_p = _lib.hostvec(dev["lo"])
            hi_keep, hi_p = _lib.hostvec(dev["hi"])
            cen_keep, cen_p = _lib.hostvec(dev["center"])
            perm_q = _lib.empty(M, t.int32)
            scr = _lib.empty(L.kdsb200_order_scratch_bytes(M), t.uint8)
            _lib.check(L.kdsb200_hilbert_order(_lib.ptr(d_pts), M, D, lo_p, hi_p, _lib.ptr(scr),
                                               _lib.ptr(perm_q), st))
            qT = _lib.empty(Mpad * 2 * D, t.float32)
            _lib.check(L.kdsb200_pack_queries_sorted_f32(_lib.ptr(d_pts), M, Mpad, D,
                                                         _lib.ptr(perm_q), cen_p, dev["grid"],
                                                         _lib.ptr(qT), st))
            nsplit = L.kdsb200_sorted_nsplit(N, M)
            scratch = _lib.empty(L.kdsb200_sorted_scratch_bytes(N, Mpad, nsplit), t.uint8)
            visited = _lib.empty(2, t.int64)
            _lib.check(L.kdsb200_kde_eval_sorted_f32(
                _lib.ptr(dev["packed"]), _lib.ptr(dev["tilebox"]), _lib.ptr(dev["tilecen"]),
                _lib.ptr(dev["tilehmax"]), N, D, _lib.ptr(qT), _lib.ptr(perm_q), M, Mpad,
                _KERNEL_ID[dev["kernel"]], dev["cull"], dev["cull_rel"], dev["out_scale"], nsplit,
                _lib.ptr(scratch), _lib.ptr(d_out), _lib.ptr(visited), st))
            #: device counters of the last call, summed over the (1024-query block, source
            #: split) work items: tiles loaded, and 512 x 1024 tile equivalents computed (cells
            #: of 32 sources x 256 queries / 64); executed pairs = last_tiles_visited * 512 * 1024
            self.last_tiles_loaded = visited[0]
            self.last_tiles_visited = visited[1].to(t.float64) / 64.0
            self.last_work_items_pairs = 512 * 1024
        elif dev["dtype"] == "float32_dense":
            Mpad = L.kdsb200_eval_mpad(M)
            qT = _lib.empty(Mpad * D, t.float32)
            cen_keep, cen_p = _lib.hostvec(dev["center"])
            _lib.check(L.kdsb200_pack_queries_f32(_lib.ptr(d_pts), M, Mpad, D, cen_p,
                                                  _lib.ptr(qT), _lib.stream()))
            nsplit = L.kdsb200_eval_nsplit(N, M)
            partial = _lib.empty(nsplit * Mpad, t.float64)
            _lib.check(L.kdsb200_kde_eval_kernel_f32(
                _lib.ptr(dev["packed"]), N, D, _lib.ptr(qT), M, Mpad, _KERNEL_ID[dev["kernel"]],
                dev["cutoff"], dev["out_scale"], nsplit, _lib.ptr(partial), _lib.ptr(d_out),
                _lib.stream()))
        else:
            nsplit = max(1, min(64, (4 * 148 * 128) // max(M, 1)))
            scratch = _lib.empty(2 * N + nsplit * M, t.float64)
            _lib.check(L.kdsb200_kde_eval_kernel_f64(
                _lib.ptr(dev["d_data"]), _lib.ptr(dev["d_w"]), _lib.ptr(dev["d_bw"]),
                dev["bw_scalar"], N, D, dev["w_scale"], _lib.ptr(d_pts), M,
                _KERNEL_ID[dev["kernel"]], dev["cutoff"], nsplit, _lib.ptr(scratch),
                _lib.ptr(d_out), _lib.stream()))
        return d_out


# KDEpy exposes the same estimator under two names; the dense sum is what
# NaiveKDE computes, TreeKDE is the name the reference imports.
NaiveKDE = TreeKDE
