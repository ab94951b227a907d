"""KDSource object: KDE particle source on an MCPL list, GPU evaluated.

API-compatible with the hot-path part of the reference's
``python/kdsource/kdsource.py`` (:32-321): ``KDSource(plist, geom, bw, J,
kernel)``, ``fit``, ``evaluate``, ``save`` and the module function ``load``,
with the same XML source-file format.  Kernel sums run in the CUDA kernels
behind :class:`kdsource_b200.treekde.TreeKDE`; bandwidth selection in
:mod:`kdsource_b200.kde`.  The ``plot_*`` methods (reference :397-923) compute
their marginal densities with the same kernels and return ``[fig, [scores,
errs]]``; the figure is drawn only when matplotlib is installed (``fig`` is None
otherwise).
"""

import ctypes
import os
from xml.dom import minidom
from xml.etree.ElementTree import Element, SubElement, parse, tostring

import numpy as np

from .geom import Geometry
from .kde import bw_methods, bw_silv, optimize_bw
from .plist import PList
from .treekde import TreeKDE

# Roughness of the kernels (reference kdsource.py:19-23)
R = {"g": 1 / (2 * np.sqrt(np.pi)), "e": 3 / 5, "b": 1 / 3}

STD_DEFAULT = 1

varnames = ["ekin", "x", "y", "z", "dx", "dy", "dz", "t"]
varmap = {name: idx for idx, name in enumerate(varnames)}
units = ["MeV", "cm", "cm", "cm", "", "", "", "ms"]

_KERNEL_FROM_CHAR = {"g": "gaussian", "e": "epa", "b": "box"}


def load(xmlfilename, N=-1):
    """Rebuild a KDSource from its XML file and fit it on N particles of the
    list (-1: all); bandwidths and scaling are taken from the file unchanged
    (reference kdsource.py:32-76)."""
    root = parse(xmlfilename).getroot()
    J = np.double(root.find("J").text)
    kelem = root.find("kernel")
    if kelem is None:
        print("No kernel specified. Using gaussian as default.")
        kern = "gaussian"
    else:
        kern = _KERNEL_FROM_CHAR.get(kelem.text, kelem.text)
    plist = PList.load(root.find("PList"))
    geom = Geometry.load(root.find("Geom"))
    scaling = np.array(root.find("scaling").text.split(), dtype="float64")
    bwel = root.find("BW")
    if bool(int(bwel.attrib["variable"])):
        bw = np.fromfile(bwel.text, dtype="float32").astype("float64")
    else:
        bw = np.double(bwel.text)
    s = KDSource(plist, geom, bw, J=J, kernel=kern)
    s.fit(N=N, scaling=scaling)
    return s


class KDSource:
    def __init__(self, plist, geom, bw="auto", J=1.0, kernel="gaussian"):
        """KDE source over `plist` with variable treatment `geom`.

        bw: float (one bandwidth), 1-D numpy array (one per particle) or the name
        of a selection method in ``bw_methods``.  J: total current [1/s].
        kernel: 'gaussian', 'box' or 'epa' (reference kdsource.py:80-130)."""
        self.plist = plist
        self.geom = geom
        self.bw_method = None
        if isinstance(bw, str):
            self.bw_method = bw
            bw = 1.0
        elif isinstance(bw, np.ndarray) and np.ndim(bw) >= 2:
            raise ValueError("BW dimension must be < 2. Use scaling for anisotropic KDE.")
        self.kde = TreeKDE(kernel=kernel, bw=bw)
        self.scaling = None
        self.J = J
        self.kernel = kernel
        self.R = R[self.kernel[0]]
        self.fitted = False

    #: lists of at least this many records are fitted on the device by default
    DEVICE_FIT_MIN = 1 << 20

    def _fit_on_device(self, N, skip, device, kwargs):
        """Device-resident fit: explicit request, or a large list on a CUDA machine; the
        host-callable sample filters of optimize_bw (weightfun / maskfun) need the host path."""
        if device is False or "weightfun" in kwargs or "maskfun" in kwargs:
            return False
        if device is True:
            return True
        return self.plist._use_device_decoder(N, skip) and (
            N < 0 or N >= self.DEVICE_FIT_MIN)

    def _fit_device(self, N, skip, scaling, importance, bw_method, sharded, kwargs):
        """:meth:`fit` without host round trips of the sample arrays: records decoded on the
        GPU (PList.get_device), Geometry.transform by kdsb200_geom_transform, weighted
        moments (N_eff, scaling) by the reduction kernels, kNN / packing from device
        tensors.  Only O(dim) statistics, the MLCV subset and the bandwidth vector (for
        `kde.bw` and `save`) reach the host."""
        from . import _lib
        from . import dist as _dist
        from .sampler import make_geom_struct

        L = _lib.load()
        t = _lib.torch()
        rank, world = _dist.rank_world()
        D = self.geom.dim
        if sharded:
            from . import mcpl as _mcpl

            with _mcpl.MCPLFile(self.plist.filename) as f:
                avail = max(0, f.nparticles - skip)
            n_req = avail if N < 0 else min(N, avail)
            lo, hi = _dist.shard_range(n_req, rank, world)
            parts, ws = self.plist.get_device(hi - lo, skip + lo)
        else:
            parts, ws = self.plist.get_device(N, skip)
        n_loc = int(parts.shape[0])
        n_tot = n_loc
        if sharded:
            n_tot = int(round(_dist.allreduce_sum_numpy(np.array([float(n_loc)]))[0]))
        if n_tot == 0:
            raise Exception("No particles for fitting.")
        print("Using {} particles for fit.".format(n_tot))
        gs = make_geom_struct(self.geom, np.ones(D), "g")
        rot_keep, rot_p = (None, None)
        if self.geom.rot is not None:
            rot_keep, rot_p = _lib.hostvec(self.geom.rot.inv().as_matrix().reshape(-1))
        vecs = _lib.empty(max(n_loc, 1) * D, t.float64)[:n_loc * D].reshape(n_loc, D)
        if n_loc:
            _lib.check(L.kdsb200_geom_transform(_lib.ptr(parts), n_loc, ctypes.byref(gs), rot_p,
                                                None, _lib.ptr(vecs), None, _lib.stream()))
        del parts

        def moments(center, power):
            if n_loc:
                sums, sw, sw2 = _lib.dev_moments(vecs, ws, n_loc, D, center, power)
            else:
                sums, sw, sw2 = np.zeros(D), 0.0, 0.0
            tot = np.concatenate((sums, [sw, sw2]))
            if sharded:
                tot = _dist.allreduce_sum_numpy(tot)
            return tot

        tot = moments(None, 1)
        self.N_eff = tot[D] ** 2 / tot[D + 1]
        if scaling is None:
            def avg(center, power):
                m = moments(center, power)
                return m[:D] / m[D]

            scaling = self.geom.std_from_moments(avg)
            if importance is None:
                importance = [1] * D
            elif len(importance) != D:
                raise Exception("importance must have len = {}.".format(D))
            scaling /= importance
        else:
            scaling = np.array(scaling, dtype=np.float64)
        scaling[scaling == 0] = STD_DEFAULT
        self.scaling = scaling
        vecs /= t.as_tensor(scaling, dtype=t.float64, device="cuda")  # in place: KDE variables
        if sharded:  # from here on every rank holds the whole sample, on its device
            vecs = _dist.allgather_concat_tensor(vecs)
            ws = _dist.allgather_concat_tensor(ws)
        if bw_method is not None:
            if bw_method not in bw_methods:
                raise Exception("Invalid bw_method. Available: {}".format(list(bw_methods)))
            self.bw_method = bw_method
        d_bw = None
        if self.bw_method is not None:
            print("Calculating bw ... ")
            if self.bw_method == "silv":
                bw = bw_silv(D, self.N_eff)
            elif self.bw_method == "mlcv":
                bw = bw_methods["mlcv"](vecs.cpu().numpy(), weights=ws.cpu().numpy(), **kwargs)
            else:  # "knn", "auto": kNN over device tensors
                bw = bw_methods[self.bw_method](vecs, weights=ws, **kwargs)
            if t.is_tensor(bw):
                d_bw, bw = bw, bw.cpu().numpy()
            bw_opt = np.reshape(bw, (-1, 1)) * self.scaling
            print("Done\nOptimal bw ({}) = {}".format(self.bw_method, bw_opt))
            self.kde = TreeKDE(kernel=self.kernel, bw=bw)
        self.kde.fit_device(vecs, ws, d_bw=d_bw)
        self.fitted = True

    def fit(self, N=-1, skip=0, scaling=None, importance=None, bw_method=None, sharded=False,
            device=None, **kwargs):
        """Load up to N particles (after `skip`) and fit the KDE model
        (reference kdsource.py:132-213): scaling = weighted std of each KDE
        variable / importance (or the given scaling; zeros -> 1), then the
        bandwidth method, if any, on the scaled variables.

        sharded=True (with torch.distributed initialised): every rank reads and
        transforms only its contiguous share of the N particles; N_eff and the scaling
        come from all-reduced weighted moments, and the transformed samples are
        all-gathered so that every rank ends with the same fitted model as an
        unsharded fit.

        device=None (default) fits lists of 2^20 records or more on the GPU without
        host round trips of the sample arrays (see :meth:`_fit_device`); True / False
        force either path.  Both give the same scaling, N_eff and bandwidths to 1e-13."""
        from . import dist as _dist

        rank, world = _dist.rank_world()
        sharded = bool(sharded) and world > 1
        if self._fit_on_device(N, skip, device, kwargs):
            return self._fit_device(N, skip, scaling, importance, bw_method, sharded, kwargs)
        if sharded:
            from . import mcpl as _mcpl

            with _mcpl.MCPLFile(self.plist.filename) as f:
                avail = max(0, f.nparticles - skip)
            n_req = avail if N < 0 else min(N, avail)
            lo, hi = _dist.shard_range(n_req, rank, world)
            parts, ws = self.plist.get(hi - lo, skip + lo)
            N = int(round(_dist.allreduce_sum_numpy(np.array([float(len(parts))]))[0]))
        else:
            parts, ws = self.plist.get(N, skip)
            N = len(parts)
        if N == 0:
            raise Exception("No particles for fitting.")
        print("Using {} particles for fit.".format(N))
        vecs = self.geom.transform(parts)
        if sharded:
            sw, sw2 = _dist.allreduce_sum_numpy(np.array([np.sum(ws), np.sum(ws ** 2)]))
            self.N_eff = sw ** 2 / sw2
        else:
            self.N_eff = np.sum(ws) ** 2 / np.sum(ws ** 2)
        if scaling is None:
            if sharded:
                scaling = self.geom.std_sharded(vecs, ws, _dist.allreduce_sum_numpy)
            else:
                scaling = self.geom.std(vecs=vecs, weights=ws)
            if importance is None:
                importance = [1] * self.geom.dim
            elif len(importance) != self.geom.dim:
                raise Exception("importance must have len = {}.".format(self.geom.dim))
            scaling /= importance
        else:
            scaling = np.array(scaling)
        scaling[scaling == 0] = STD_DEFAULT
        self.scaling = scaling
        if sharded:  # from here on every rank holds the whole (transformed) sample
            vecs = _dist.allgather_rows(vecs)
            ws = _dist.allgather_rows(ws)
        if bw_method is not None:
            if bw_method not in bw_methods:
                raise Exception("Invalid bw_method. Available: {}".format(list(bw_methods)))
            self.bw_method = bw_method
        if self.bw_method is not None:
            print("Calculating bw ... ")
            bw = optimize_bw(self.bw_method, vecs / self.scaling, ws, **kwargs)
            bw_opt = np.reshape(bw, (-1, 1)) * self.scaling
            print("Done\nOptimal bw ({}) = {}".format(self.bw_method, bw_opt))
            self.kde = TreeKDE(kernel=self.kernel, bw=bw)
        self.kde.fit(vecs / self.scaling, weights=ws)
        self.fitted = True

    def evaluate(self, parts):
        """Density and statistical error at `parts` ((obs, 8) columns
        [ekin,x,y,z,dx,dy,dz,t]); returns ``[evals, errs]``
        (reference kdsource.py:215-248)."""
        if self.fitted is False:
            raise Exception("Must fit before evaluating.")
        parts = np.ascontiguousarray(parts, dtype=np.float64)
        if parts.ndim != 2 or parts.shape[1] not in (7, 8):
            raise ValueError("parts must have shape (obs, 8)")
        if parts.shape[1] == 7:  # no time column
            parts = np.concatenate((parts, np.zeros((len(parts), 1))), axis=1)
        vol = np.prod(self.scaling)
        dens, jacs = self._density_and_jac(parts)
        evals = 1 / vol * dens
        errs = np.sqrt(
            evals * self.R ** self.geom.dim / (self.N_eff * np.mean(self.kde.bw) * vol))
        evals *= self.J * jacs
        errs *= self.J * jacs
        return [evals, errs]

    def _density_and_jac(self, parts, chunk=1 << 22):
        """kde.evaluate(geom.transform(parts) / scaling) and geom.jac(parts), with
        the transform, the scaling and the kernel sum all on the device
        (kdsb200_geom_transform + the evaluation kernel); `parts` is not modified."""
        from . import _lib
        from .sampler import make_geom_struct

        L = _lib.load()
        t = _lib.torch()
        D = self.geom.dim
        gs = make_geom_struct(self.geom, np.ones(D), "g")
        inv_keep, inv_p = _lib.hostvec(1.0 / np.asarray(self.scaling, dtype=np.float64))
        rot_keep, rot_p = (None, None)
        if self.geom.rot is not None:
            rot_keep, rot_p = _lib.hostvec(self.geom.rot.inv().as_matrix().reshape(-1))
        from . import dist as _dist

        dev = self.kde._device_state()
        M = len(parts)
        # query points are split contiguously over the ranks (sources replicated)
        rank, world = _dist.rank_world()
        lo, hi = _dist.shard_range(M, rank, world)
        def work(d_parts, n):
            d_vecs = _lib.empty(n * D, t.float64)
            d_jac = _lib.empty(n, t.float64)
            _lib.check(L.kdsb200_geom_transform(_lib.ptr(d_parts), n, ctypes.byref(gs), rot_p,
                                                inv_p, _lib.ptr(d_vecs), _lib.ptr(d_jac),
                                                _lib.stream()))
            return self.kde.evaluate_device(d_vecs, n, dev), d_jac

        n_loc = hi - lo
        chunk = int(min(chunk, max(1 << 16, -(-n_loc // 4))))
        (dens, jacs), _ = _lib.stream_pieces(parts, lo, hi, chunk, work, n_out=2, tag="kds_eval")
        if world > 1:  # the ranks' disjoint slices, gathered in rank (= query) order
            dens = _dist.allgather_rows(dens)
            jacs = _dist.allgather_rows(jacs)
        return dens, jacs

    def save(self, xmlfilename=None, bwfile=None, adjust_N=True):
        """Write the XML source file (and the float32 bandwidth file for adaptive
        bandwidths); returns the XML file name (reference kdsource.py:250-321).

        adjust_N: rescale the bandwidth by bw_silv(dim, N_tot)/bw_silv(dim, N_fit)
        so a fit on a subset applies to the whole list."""
        stem = self.plist.filename.split(".")[0]
        if xmlfilename is None:
            xmlfilename = stem + "_source.xml"
        if bwfile is None:
            bwfile = bwfilename = stem + "_bws"
        elif isinstance(bwfile, str):
            bwfilename = bwfile
        else:  # file object
            bwfilename = bwfile.name
        bw = self.kde.bw
        if adjust_N:
            dim = self.geom.dim
            if not self.plist.params_set:
                self.plist.set_params()
            factor = bw_silv(dim, self.plist.N_eff) / bw_silv(dim, self.N_eff)
            if np.isscalar(bw):
                # local only: `bw *= factor` on a float rebinds the name in the reference
                # (kdsource.py:293), so self.kde.bw keeps its value
                bw = bw * factor
            else:
                bw *= factor  # in place, like the reference (kdsource.py:293)
        root = Element("KDSource")
        Jel = SubElement(root, "J")
        Jel.set("units", "1/s")
        Jel.text = str(self.J)
        SubElement(root, "kernel").text = str(self.kernel[0])
        self.plist.save(SubElement(root, "PList"))
        self.geom.save(SubElement(root, "Geom"))
        SubElement(root, "scaling").text = np.array_str(self.scaling)[1:-1]
        bwel = SubElement(root, "BW")
        if np.isscalar(bw):
            bwel.set("variable", "0")
            bwel.text = str(bw)
        else:
            bwel.set("variable", "1")
            bw.astype("float32").tofile(bwfile, format="float32")
            print("Bandwidth file: {}".format(bwfilename))
            bwel.text = os.path.abspath(bwfilename)
        xmlstr = minidom.parseString(tostring(root, encoding="utf8", method="xml")).toprettyxml()
        with open(xmlfilename, "w") as file:
            file.write(xmlstr)
        print("Successfully saved parameters file {}".format(xmlfilename))
        return xmlfilename

    # ------------------------------------------------------------------
    # plot back-ends (reference kdsource.py:397-923): a marginal KDE over a column
    # subset of the fitted samples inside [vec0, vec1], evaluated on a grid
    # ------------------------------------------------------------------
    def _range_mask(self, vec0, vec1):
        data = self.kde.data
        mask = np.ones(len(data), dtype=bool)
        if vec0 is not None:
            mask &= np.logical_and.reduce(np.asarray(vec0) / self.scaling <= data, axis=1)
        if vec1 is not None:
            mask &= np.logical_and.reduce(data <= np.asarray(vec1) / self.scaling, axis=1)
        return mask

    def _marginal(self, cols, points, vec0, vec1, adjust_bw, r_power):
        """Scores and errors of the marginal KDE over columns `cols` at the (already
        parametrised, unscaled) `points` ((M, len(cols))).  Same arithmetic as the
        reference's plot methods; two deliberate differences: a per-particle
        bandwidth array is masked together with the samples (the reference passes
        the unmasked array), and ``adjust_bw`` does not modify ``self.kde.bw``."""
        if self.fitted is False:
            raise Exception("Must fit before plotting.")
        mask = self._range_mask(vec0, vec1)
        if np.sum(mask) == 0:
            raise Exception("No particles in [vec0,vec1] range.")
        cols = np.atleast_1d(cols)
        vecs = self.kde.data[:, cols][mask]
        allw = self.kde.weights if self.kde.weights is not None else np.ones(len(mask))
        ws = allw[mask]
        N_eff = np.sum(ws) ** 2 / np.sum(ws ** 2)
        scaling = self.scaling[cols]
        bw = self.kde.bw
        if np.ndim(bw) == 1:
            bw = np.asarray(bw)[:len(mask)][mask]
        if adjust_bw:
            bw = bw * (bw_silv(1, N_eff) / bw_silv(self.geom.dim, self.N_eff))
        kde = TreeKDE(kernel=self.kernel, bw=bw)
        kde.fit(vecs, weights=ws)
        vol = np.prod(scaling)
        scores = 1 / vol * kde.evaluate(np.asarray(points, dtype=np.float64) / scaling)
        errs = np.sqrt(scores * self.R ** r_power / (N_eff * np.mean(bw) * vol))
        frac = self.J * np.sum(ws) / np.sum(allw)
        return scores * frac, errs * frac

    @staticmethod
    def _pyplot():
        try:
            import matplotlib.pyplot as plt
        except ImportError:
            return None
        return plt

    def _errorbar(self, x, scores, errs, label, xscale, yscale, xlabel, ylabel):
        plt = self._pyplot()
        if plt is None:
            return None
        plt.errorbar(x, scores, errs, label=label, capsize=1, linewidth=1)
        plt.xscale(xscale)
        plt.yscale(yscale)
        plt.xlabel(xlabel)
        plt.ylabel(ylabel)
        plt.grid()
        plt.legend()
        plt.tight_layout()
        return plt.gcf()

    def _colormesh(self, grids, scores, scale, title, xlabel, ylabel):
        plt = self._pyplot()
        if plt is None:
            return None
        edges = [np.concatenate((g[:1], (g[1:] + g[:-1]) / 2, g[-1:])) for g in grids]
        norm = None
        if scale == "log":
            import matplotlib.colors as col

            norm = col.LogNorm()
        plt.pcolormesh(edges[0], edges[1], scores.reshape(len(grids[1]), len(grids[0])),
                       cmap="jet", norm=norm, rasterized=True)
        plt.colorbar()
        plt.title(title)
        plt.xlabel(xlabel)
        plt.ylabel(ylabel)
        plt.tight_layout()
        return plt.gcf()

    def plot_integr(self, var, grid, vec0=None, vec1=None, **kwargs):
        """Univariate distribution along the parametrised variable `var` (name or
        index), integrated over [vec0, vec1] in the other variables
        (reference kdsource.py:397-508).  kwargs: xscale, yscale, fact, label,
        adjust_bw.  Returns ``[fig, [scores, errs]]``."""
        if self.fitted is False:
            raise Exception("Must fit before plotting.")
        if isinstance(var, str):
            var = self.geom.varmap[var]
        grid = np.asarray(grid, dtype=np.float64)
        scores, errs = self._marginal([var], grid.reshape(-1, 1), vec0, vec1,
                                      kwargs.get("adjust_bw", False), 1)
        if "fact" in kwargs:
            scores *= kwargs["fact"]
            errs *= kwargs["fact"]
        fig = self._errorbar(
            grid, scores, errs, kwargs.get("label", "KDE"), kwargs.get("xscale", "linear"),
            kwargs.get("yscale", "log"),
            r"${}\ [{}]$".format(self.geom.varnames[var], self.geom.units[var]),
            r"$J\ \left[ \frac{{{}}}{{{}\ s}} \right]$".format(self.plist.pt, self.geom.units[var]))
        return [fig, [scores, errs]]

    def _plot_metric_axis(self, which, grid_x, vec0, vec1, kwargs, xlabel, ylabel):
        """plot_E / plot_t: the first (energy) or last (time) metric's variable, with
        the metric's jacobian applied so the spectrum is per MeV / per ms."""
        if self.fitted is False:
            raise Exception("Must fit before plotting.")
        col = 0 if which == 0 else self.geom.dim - 1
        metric = self.geom.ms[which]
        grid_x = np.asarray(grid_x, dtype=np.float64)
        pts = np.asarray(metric.transform(grid_x), dtype=np.float64).reshape(-1, 1)
        jacs = np.asarray(metric.jac(grid_x), dtype=np.float64).reshape(-1)
        scores, errs = self._marginal([col], pts, vec0, vec1, kwargs.get("adjust_bw", False), 1)
        scores *= jacs
        errs *= jacs
        if "fact" in kwargs:
            scores *= kwargs["fact"]
            errs *= kwargs["fact"]
        fig = self._errorbar(grid_x, scores, errs, kwargs.get("label", "KDE"), "log", "log",
                             xlabel, ylabel)
        return [fig, [scores, errs]]

    def plot_E(self, grid_E, vec0=None, vec1=None, **kwargs):
        """Energy spectrum [1/(MeV s)]; assumes energy is the first metric
        (reference kdsource.py:510-608).  kwargs: fact, label, adjust_bw."""
        return self._plot_metric_axis(
            0, grid_E, vec0, vec1, kwargs, r"$E\ [MeV]$",
            r"$J\ \left[ \frac{{{}}}{{MeV\ s}} \right]$".format(self.plist.pt))

    def plot_t(self, grid_t, vec0=None, vec1=None, **kwargs):
        """Time distribution; assumes time is the last metric
        (reference kdsource.py:610-708).  kwargs: fact, label, adjust_bw."""
        return self._plot_metric_axis(
            -1, grid_t, vec0, vec1, kwargs, r"$t$ [ms]",
            r"$J\ \left[ \frac{{{}}}{{s^2}} \right]$".format(self.plist.pt))

    def plot_point(self, var, grid, part0, **kwargs):
        """Full multivariate density along one particle variable, the others fixed at
        `part0` (reference kdsource.py:324-395).  kwargs: xscale ('linear'), yscale
        ('log'), fact, label.  Returns [fig, [scores, errs]]; `part0` is not modified
        (the reference zeroes its `var` entry in place)."""
        if self.fitted is False:
            raise Exception("Must fit before plotting.")
        if isinstance(var, str):
            var = varmap[var]
        grid = np.asarray(grid, dtype=np.float64)
        parts = np.zeros((len(grid), 7))
        parts[:, var] = grid
        base = np.array(part0, dtype=np.float64)
        base[var] = 0
        parts += base
        scores, errs = self.evaluate(parts)
        if "fact" in kwargs:
            scores *= kwargs["fact"]
            errs *= kwargs["fact"]
        fig = self._errorbar(
            grid, scores, errs, kwargs.get("label", "KDE"), kwargs.get("xscale", "linear"),
            kwargs.get("yscale", "log"), r"${}\ [{}]$".format(varnames[var], units[var]),
            r"$J\ \left[ \frac{{{}}}{{{} s}} \right]$".format(self.plist.pt, self.geom.volunits))
        return [fig, [scores, errs]]

    def plot2D_point(self, vrs, grids, part0, **kwargs):
        """Full multivariate density on a 2-D grid of two particle variables, the
        others fixed at `part0` (reference kdsource.py:710-785).  The reference
        multiplies the list returned by ``evaluate`` by J (a TypeError for float J);
        ``evaluate`` already includes J, so it is used as is."""
        if self.fitted is False:
            raise Exception("Must fit before plotting.")
        vrs = [varmap[v] if isinstance(v, str) else v for v in vrs]
        grids = [np.asarray(g, dtype=np.float64) for g in grids]
        parts = np.zeros((len(grids[0]) * len(grids[1]), 7))
        parts[:, vrs] = np.reshape(np.meshgrid(*grids), (2, -1)).T
        part0 = np.array(part0, dtype=np.float64)
        part0[vrs] = 0
        parts += part0
        scores, errs = self.evaluate(parts)
        if "fact" in kwargs:
            scores *= kwargs["fact"]
            errs *= kwargs["fact"]
        title = kwargs.get("title", "KDE") + "\n" + (
            r"$J\ \left[ \frac{{{}}}{{{}\ s}} \right]$".format(self.plist.pt, self.geom.volunits))
        fig = self._colormesh(grids, scores, kwargs.get("scale", "linear"), title,
                              r"${}\ [{}]$".format(varnames[vrs[0]], units[vrs[0]]),
                              r"${}\ [{}]$".format(varnames[vrs[1]], units[vrs[1]]))
        return [fig, [scores, errs]]

    def plot2D_integr(self, vrs, grids, vec0=None, vec1=None, **kwargs):
        """Bivariate distribution along two parametrised variables, integrated over
        [vec0, vec1] in the others (reference kdsource.py:787-923).  kwargs: scale,
        fact, title, adjust_bw."""
        if self.fitted is False:
            raise Exception("Must fit before plotting.")
        vrs = [self.geom.varmap[v] if isinstance(v, str) else v for v in vrs]
        grids = [np.asarray(g, dtype=np.float64) for g in grids]
        pts = np.reshape(np.meshgrid(*grids), (2, -1)).T
        scores, errs = self._marginal(vrs, pts, vec0, vec1, kwargs.get("adjust_bw", False), 2)
        if "fact" in kwargs:
            scores *= kwargs["fact"]
            errs *= kwargs["fact"]
        u0, u1 = self.geom.units[vrs[0]], self.geom.units[vrs[1]]
        un = u0 + "^2" if u0 == u1 else u0 + u1
        title = kwargs.get("title", "KDE") + "\n" + (
            r"$J\ \left[ \frac{{{}}}{{{}\ s}} \right]$".format(self.plist.pt, un))
        fig = self._colormesh(grids, scores, kwargs.get("scale", "linear"), title,
                              r"${}\ [{}]$".format(self.geom.varnames[vrs[0]], u0),
                              r"${}\ [{}]$".format(self.geom.varnames[vrs[1]], u1))
        return [fig, [scores, errs]]
