"""Device resampler: weighted pick + kernel perturbation + MCPL-3 packing.

Host-side driver of ``kdsb200_weights_cdf`` / ``kdsb200_resample``.  It plays
the role of the reference's ``KDSource``/``PList``/``Geometry`` C structs
(clib/include/kdsource/*.h) for the sampling loop of
``clib/src/resamplemcpl.c:35-43``: the valid particles of the list (after the
PList translation/rotation/x2z of ``PList_get``, plist.c:69-89), their
per-particle bandwidths and the metric descriptors live in HBM, and output
particles are produced in batches, each identified by its global index so a
run can be split over calls, streams or GPUs without changing the result.
"""

import ctypes

import numpy as np

from . import _lib
from .geom import METRIC_ORDER
from .utils import pt2pdg

KERNEL_CHAR = {"gaussian": b"g", "epa": b"e", "box": b"b", "g": b"g", "e": b"e", "b": b"b"}
DEFAULT_SLOTS = 32


def make_geom_struct(geom, scaling, kernel="g"):
    """``kdsb200_geom`` for a :class:`Geometry`, the flat `scaling` vector (one
    entry per KDE variable, cast to float as Metric_create does, geom.c:35-38)
    and a kernel name/char."""
    if len(geom.ms) > _lib.MAX_METRICS:
        raise ValueError("at most {} metrics per geometry".format(_lib.MAX_METRICS))
    scaling = np.asarray(scaling, dtype=np.float64).reshape(-1)
    if len(scaling) != geom.dim:
        raise ValueError("scaling must have len = {}".format(geom.dim))
    g = _lib.Geom()
    g.order = len(geom.ms)
    col = 0
    for i, m in enumerate(geom.ms):
        name = m.__class__.__name__
        if name not in METRIC_ORDER:
            raise ValueError("metric {} has no device perturbation".format(name))
        d = g.ms[i]
        d.type = METRIC_ORDER.index(name)
        d.dim = m.dim
        for j in range(m.dim):
            d.scaling[j] = np.float32(scaling[col + j])
        col += m.dim
        pars = m.c_params()
        d.nps = len(pars)
        for j, v in enumerate(pars):
            d.params[j] = v
    if kernel not in KERNEL_CHAR:
        raise ValueError("unknown kernel {!r}".format(kernel))
    g.kernel = KERNEL_CHAR[kernel]
    g.has_trasl = int(geom.trasl is not None)
    g.has_rot = int(geom.rot is not None)
    if geom.trasl is not None:
        for j in range(3):
            g.trasl[j] = float(geom.trasl[j])
    if geom.rot is not None:
        rv = geom.rotvec()
        for j in range(3):
            g.rot[j] = float(rv[j])
    return g


class DeviceSampler:
    """Resampler over one particle list resident on the current CUDA device.

    ``precision="float32"`` (default) keeps the list as fp32 and perturbs in fp32 for the
    two geometries with specialised kernels (every coordinate within 1e-5 of its scale);
    any other geometry, and ``precision="float64"``, use the fp64 store (1e-12)."""

    def __init__(self, parts, weights, bw, geom_struct, pdgcode=2112, precision="float32",
                 contract=2, pick_table="auto", pick_bits=None):
        L = _lib.load()
        t = _lib.torch()
        on_device = t.is_tensor(parts)  # (N,8) / (N,) float64 CUDA tensors (PList.get_device)
        if on_device:
            parts = parts.to(device="cuda", dtype=t.float64).contiguous()
            weights = weights.to(device="cuda", dtype=t.float64).contiguous()
        else:
            parts = np.ascontiguousarray(parts, dtype=np.float64)
            weights = np.ascontiguousarray(weights, dtype=np.float64)
        if parts.ndim != 2 or parts.shape[1] != 8 or len(parts) == 0:
            raise ValueError("parts must have shape (N, 8) with N > 0")
        if len(weights) != len(parts):
            raise ValueError("weights must have one entry per particle")
        if bool((weights <= 0).any()):
            raise ValueError("weights must be positive (filter w <= 0 first)")
        if precision not in ("float32", "float64"):
            raise ValueError("precision must be 'float32' or 'float64'")
        if contract not in (1, 2):
            raise ValueError("contract must be 1 (polar method, as the reference) or 2 "
                             "(Box-Muller pairs)")
        #: uniform-consumption contract of the Philox mode (kdsb200.h, kdsb200_resample)
        self.contract = int(contract)
        self.N = len(parts)
        self.precision = precision
        self.geom = geom_struct
        self.pdgcode = int(pdgcode)
        # fp64 sum in a fixed order on the host (the CDF scale must not depend on where
        # the list was decoded)
        self.sum_w = float(np.sum(weights.cpu().numpy() if on_device else weights))
        d_parts = parts if on_device else _lib.to_dev(parts)
        # fp32 store + arithmetic only for the geometries with a specialised kernel; the
        # generic maps (acos / atan2 based metrics, rotations) run on the fp64 store, where
        # fp32-rounded sources could not hold 1e-5 of the variable's scale
        self.f32 = precision == "float32" and bool(
            L.kdsb200_resample_fp32_ok(ctypes.byref(geom_struct)))
        self.src = _lib.empty(self.N * 8, t.float32 if self.f32 else t.float64)
        _lib.check(L.kdsb200_convert_particles(_lib.ptr(d_parts), self.N, int(self.f32),
                                               _lib.ptr(self.src), _lib.stream()))
        d_w = weights if on_device else _lib.to_dev(weights)
        scratch = _lib.empty(L.kdsb200_cdf_scratch_words(self.N), t.int64)
        self.cdf = _lib.empty(self.N, t.int64)  # uint64 payload
        _lib.check(L.kdsb200_weights_cdf(_lib.ptr(d_w), self.N, self.sum_w, _lib.ptr(scratch),
                                         _lib.ptr(self.cdf), _lib.stream()))
        if np.ndim(bw) == 0:
            self.bw = None
            self.bw_scalar = float(bw)
        else:
            bw = np.ascontiguousarray(bw, dtype=np.float32).reshape(-1)
            if len(bw) < self.N:
                raise ValueError("bandwidth array shorter than the particle list")
            self.bw = _lib.to_dev(bw[:self.N], np.float32)
            self.bw_scalar = 0.0
        self.guide, self.guide_bits, self.guide_kind = None, 0, 0
        self.pick, self.pick_bits = None, 0
        if self.N < (1 << 32):
            # search window table: one CDF entry per bucket while the table stays L2-resident
            # (measured +7 % at N=1e5), ~8 per bucket for large lists where the table itself
            # would become a second random HBM access; at most 2^24 buckets
            per_bucket = 1.0 if self.N <= (1 << 21) else 8.0
            self.guide_bits = int(min(24, max(4, np.ceil(np.log2(max(self.N, 16) / per_bucket)))))
            self.guide = _lib.empty((1 << self.guide_bits) + 1, t.int32)
            _lib.check(L.kdsb200_cdf_guide(_lib.ptr(self.cdf), self.N, self.guide_bits,
                                           _lib.ptr(self.guide), _lib.stream()))
            # pick table for the fp32 Philox kernels (kdsb200_pick_table): one 128-byte line
            # per output particle in place of the dependent guide -> CDF probes -> particle ->
            # bandwidth chain (five lines): 2x the rate on lists that have outgrown the L2,
            # +4..7 % on short ones.  The table takes 256-512 B per source.
            if pick_table == "auto":
                pick_table = self.f32
            if pick_table:
                if not self.f32:
                    raise ValueError("the pick table needs the fp32 store of a specialised geometry")
                # >= 2 buckets per source: ~1 pick in 30 then goes beyond the two inlined
                # particles; one bucket fewer per halving of what the free memory allows
                bits = int(np.ceil(np.log2(2.0 * max(self.N, 8))))
                free = t.cuda.mem_get_info()[0]
                while bits > 4 and int(L.kdsb200_pick_table_bytes(bits)) > free // 3:
                    bits -= 1
                if pick_bits is not None:
                    bits = int(pick_bits)
                try:
                    self.pick = _lib.empty(int(L.kdsb200_pick_table_bytes(bits)), t.uint8)
                except t.cuda.OutOfMemoryError:  # the window table serves every kernel
                    self.pick = None
                if self.pick is not None:
                    self.pick_bits = bits
                    _lib.check(L.kdsb200_pick_table(_lib.ptr(self.cdf), self.N, bits,
                                                    _lib.ptr(self.src), _lib.ptr(self.bw),
                                                    self.bw_scalar, _lib.ptr(self.pick),
                                                    _lib.stream()))
        t.cuda.current_stream().synchronize()

    def sample_device(self, count, seed=0, first=0, ext_uniforms=None, perturb=True,
                      want_records=True, want_idx=False, want_parts=False, out=None,
                      seq_start=-1):
        """Launch the resample kernel for output indices first..first+count-1.
        Returns a dict of device tensors ("records" uint8 (count*36), "idx" int64,
        "parts" float64 (count*8)).  `out` may supply a preallocated records
        tensor."""
        L = _lib.load()
        t = _lib.torch()
        count = int(count)
        res = {}
        d_ext, slots = None, 0
        if ext_uniforms is not None:
            ext = np.ascontiguousarray(ext_uniforms, dtype=np.float64)
            if ext.ndim != 2 or ext.shape[0] < count:
                raise ValueError("ext_uniforms must have shape (count, slots)")
            slots = ext.shape[1]
            d_ext = _lib.to_dev(ext[:count])
        if want_records:
            res["records"] = out if out is not None else _lib.empty(count * 36, t.uint8)
        if want_idx:
            res["idx"] = _lib.empty(count, t.int64)
        if want_parts:
            res["parts"] = _lib.empty(count * 8, t.float64)
        # the pick table serves the Philox mode; external uniforms go through the window table
        use_pick = self.pick is not None and d_ext is None
        _lib.check(L.kdsb200_resample(
            _lib.ptr(self.src), int(self.f32), _lib.ptr(self.cdf),
            _lib.ptr(self.pick if use_pick else self.guide),
            self.pick_bits if use_pick else self.guide_bits, int(use_pick), _lib.ptr(self.bw),
            self.bw_scalar, self.N, ctypes.byref(self.geom), self.pdgcode,
            int(seed) & 0xFFFFFFFFFFFFFFFF, int(first), count, int(seq_start), _lib.ptr(d_ext),
            slots,
            int(bool(perturb)), self.contract, _lib.ptr(res.get("records")),
            _lib.ptr(res.get("idx")),
            _lib.ptr(res.get("parts")), _lib.stream()))
        return res

    def sample(self, count, seed=0, first=0, ext_uniforms=None, perturb=True,
               want=("records",)):
        """Host-returning convenience wrapper around :meth:`sample_device`."""
        r = self.sample_device(count, seed, first, ext_uniforms, perturb,
                               "records" in want, "idx" in want, "parts" in want)
        out = {}
        if "records" in r:
            out["records"] = r["records"].cpu().numpy().reshape(count, 36)
        if "idx" in r:
            out["idx"] = r["idx"].cpu().numpy()
        if "parts" in r:
            out["parts"] = r["parts"].cpu().numpy().reshape(count, 8)
        return out

    def cdf_host(self):
        return self.cdf.cpu().numpy().view(np.uint64)
