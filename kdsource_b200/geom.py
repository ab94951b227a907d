"""Geometry and Metric objects: particle variables <-> KDE variables.

API-compatible with the reference's ``python/kdsource/geom.py`` (same class
names, constructor arguments, ``transform`` / ``inverse_transform`` / ``jac`` /
``mean`` / ``std`` / ``save`` / ``load`` and the ``_metrics`` registry with the
``Geom*`` factories), written table-driven: each metric declares its particle
columns, variable names and XML parameter list once, and the XML and C-struct
serialisation is generic.

Particle arrays have columns [ekin, x, y, z, dx, dy, dz, t]
(kdsource.py:27).  Metric ids follow ``_metric_names`` in clib/src/geom.c:13-16
so the same geometry can be handed to the device sampler.
"""

from xml.etree.ElementTree import SubElement

import numpy as np

METRIC_ORDER = [
    "Energy", "Lethargy", "Wavelength", "Vol", "SurfXY", "SurfR", "SurfR2",
    "SurfCircle", "Guide", "Isotrop", "Polar", "PolarMu", "Time", "Decade",
]

_DEG = 180.0 / np.pi


def _fmt(v):
    """Parameter formatting of the reference writers ("{}".format)."""
    return "{}".format(v)


class Metric:
    """Transformation of a subset of particle variables (reference geom.py:10-105)."""

    #: names of the constructor parameters serialised into <params>
    _param_names = ()
    #: parameters parsed as integers on load (Isotrop keep flags)
    _int_params = False

    def __init__(self, partvars, varnames, units, volunits):
        if len(varnames) != len(units):
            raise ValueError("varnames and units must have same len.")
        self.partvars = partvars
        self.dim = len(varnames)
        self.varnames = varnames
        self.varmap = {name: idx for idx, name in enumerate(varnames)}
        self.units = units
        self.volunits = volunits

    # -- maps (identity by default) -----------------------------------------
    def transform(self, parts):
        return parts

    def inverse_transform(self, vecs):
        return vecs

    def jac(self, parts):
        return np.ones(len(parts))

    # -- statistics in parametrised space ------------------------------------
    def mean(self, parts=None, vecs=None, weights=None):
        if vecs is None:
            vecs = self.transform(parts)
        return self.inverse_transform(np.average(vecs, axis=0, weights=weights))

    def std(self, parts=None, vecs=None, weights=None):
        if vecs is None:
            vecs = self.transform(parts)
        centre = np.average(vecs, axis=0, weights=weights)
        return np.sqrt(np.average((vecs - centre) ** 2, axis=0, weights=weights))

    # -- the same statistics when every rank holds a share of the list:
    # `allsum` adds a numpy array over the ranks.  Two phases, like the formulas
    # above: weighted mean first, then the weighted mean squared deviation from it.
    @staticmethod
    def _wavg_sharded(vals, weights, allsum):
        if weights is None:
            num, den = vals.sum(axis=0), np.array([float(len(vals))])
        else:
            num, den = (vals * weights[:, None]).sum(axis=0), np.array([weights.sum()])
        tot = allsum(np.concatenate((num, den)))
        return tot[:-1] / tot[-1]

    def mean_sharded(self, vecs, weights, allsum):
        return self._wavg_sharded(vecs, weights, allsum)

    def std_sharded(self, vecs, weights, allsum):
        centre = self.mean_sharded(vecs, weights, allsum)
        return np.sqrt(self._wavg_sharded((vecs - centre) ** 2, weights, allsum))

    # -- the same statistics from column moments computed elsewhere (on the device):
    # `avg(centre, power)` returns the weighted averages of (v - centre)^power over this
    # metric's columns (centre None = 0), e.g. from kdsb200_weighted_moments.
    def std_from_moments(self, avg):
        return np.sqrt(avg(avg(None, 1), 2))

    # -- serialisation --------------------------------------------------------
    def params(self):
        return [getattr(self, n) for n in self._param_names]

    def save(self, mtree):
        SubElement(mtree, "dim").text = str(self.dim)
        pel = SubElement(mtree, "params")
        vals = self.params()
        pel.set("nps", str(len(vals)))
        if vals:
            if self._int_params:
                pel.text = " ".join("{:d}".format(int(v)) for v in vals)
            else:
                pel.text = " ".join(_fmt(v) for v in vals)

    @classmethod
    def load(cls, mtree):
        nps = len(cls._param_names)
        if nps == 0:
            return cls()
        dim = int(mtree[0].text)
        vals = np.array(mtree[1].text.split(), dtype="int" if cls._int_params else "float64")
        proto_dim = cls._xml_dim
        if dim != proto_dim or len(vals) != nps or int(mtree[1].attrib["nps"]) != nps:
            raise Exception("Invalid metric tree.")
        return cls(*vals)

    def c_params(self):
        """Parameter vector in the order the C perturbation functions read it."""
        out = []
        for v in self.params():
            out.append(0.0 if v is None else float(v))
        return out


class Geometry(Metric):
    """Full particle-variable treatment (reference geom.py:108-294)."""

    def __init__(self, metrics, trasl=None, rot=None):
        varnames = sum([m.varnames for m in metrics], [])
        units = sum([m.units for m in metrics], [])
        volunits = " ".join(m.volunits for m in metrics)
        super().__init__(range(8), varnames, units, volunits)
        self.ms = metrics
        if trasl is not None:
            trasl = np.array(trasl).reshape(-1)
            if trasl.shape != (3,):
                raise ValueError("Invalid trasl.")
        self.trasl = trasl
        self.rot = _as_rotation(rot, "Invalid rot.")

    def _to_local(self, parts):
        # like the reference (geom.py:160-168, 193-201) this edits `parts` in place
        if self.trasl is not None:
            parts[:, 1:4] -= self.trasl
        if self.rot is not None:
            parts[:, 1:4] = self.rot.apply(parts[:, 1:4], inverse=True)
            parts[:, 4:7] = self.rot.apply(parts[:, 4:7], inverse=True)

    def transform(self, parts):
        self._to_local(parts)
        # written block by block into a C-ordered array: np.concatenate would inherit the
        # column-major layout of the fancy-indexed inputs, and every consumer (scaling,
        # uploads, the KDE kernels' staging) would then pay a strided copy
        blocks = [m.transform(parts[:, m.partvars]) for m in self.ms]
        out = np.empty((len(parts), sum(b.shape[1] for b in blocks)),
                       dtype=np.result_type(*blocks) if blocks else np.float64)
        col = 0
        for b in blocks:
            out[:, col:col + b.shape[1]] = b
            col += b.shape[1]
        return out

    def inverse_transform(self, vecs):
        parts = np.zeros((len(vecs), 8))
        col = 0
        for m in self.ms:
            parts[:, m.partvars] = m.inverse_transform(vecs[:, col:col + m.dim])
            col += m.dim
        if self.trasl is not None:
            parts[:, 1:4] += self.trasl
        if self.rot is not None:
            parts[:, 1:4] = self.rot.apply(parts[:, 1:4])
            parts[:, 4:7] = self.rot.apply(parts[:, 4:7])
        return parts

    def jac(self, parts):
        self._to_local(parts)
        # product over metrics, one value per particle.  (The reference reduces over
        # axis=1, geom.py:205, i.e. over particles, which only broadcasts in
        # evaluate() for single-particle input; the per-particle product is what
        # kdsource.py:246-247 needs.)
        return np.prod([m.jac(parts[:, m.partvars]) for m in self.ms], axis=0)

    def _per_metric(self, fn, parts, vecs, weights):
        if vecs is None:
            vecs = self.transform(parts)
        out, col = [], 0
        for m in self.ms:
            out.append(getattr(m, fn)(vecs=vecs[:, col:col + m.dim], weights=weights))
            col += m.dim
        return np.concatenate(out)

    def mean(self, parts=None, vecs=None, weights=None):
        return self._per_metric("mean", parts, vecs, weights)

    def std(self, parts=None, vecs=None, weights=None):
        return self._per_metric("std", parts, vecs, weights)

    def std_sharded(self, vecs, weights, allsum):
        """:meth:`std` of the union of the ranks' samples (`vecs`, `weights` are this
        rank's share; `allsum` sums an array over the ranks)."""
        out, col = [], 0
        for m in self.ms:
            out.append(m.std_sharded(vecs[:, col:col + m.dim], weights, allsum))
            col += m.dim
        return np.concatenate(out)

    def std_from_moments(self, avg):
        """:meth:`std` from a column-moment callback over all KDE variables:
        `avg(centre (dim,) or None, power)` -> (dim,) weighted averages."""
        out, col = [], 0
        for m in self.ms:
            lo, hi = col, col + m.dim

            def sub(centre, power, lo=lo, hi=hi):
                full = None
                if centre is not None:
                    full = np.zeros(self.dim)
                    full[lo:hi] = centre
                return avg(full, power)[lo:hi]

            out.append(m.std_from_moments(sub))
            col += m.dim
        return np.concatenate(out)

    def save(self, gtree):
        gtree.set("order", str(len(self.ms)))
        for m in self.ms:
            m.save(SubElement(gtree, m.__class__.__name__))
        SubElement(gtree, "trasl").text = (
            np.array_str(self.trasl)[1:-1] if self.trasl is not None else "")
        SubElement(gtree, "rot").text = (
            np.array_str(self.rot.as_rotvec())[1:-1] if self.rot is not None else "")

    @staticmethod
    def load(gtree):
        order = int(gtree.attrib["order"])
        metrics = []
        for i in range(order):
            name = gtree[i].tag
            if name not in _metrics:
                raise Exception("Invalid metricname {}".format(name))
            metrics.append(_metrics[name].load(gtree[i]))
        trasl = np.array(gtree[-2].text.split(), dtype="float64") if gtree[-2].text else None
        rot = np.array(gtree[-1].text.split(), dtype="float64") if gtree[-1].text else None
        return Geometry(metrics, trasl=trasl, rot=rot)

    def rotvec(self):
        return None if self.rot is None else np.asarray(self.rot.as_rotvec(), dtype=np.float64)


def _as_rotation(rot, msg):
    if rot is None:
        return None
    import scipy.spatial.transform as st

    if isinstance(rot, st.Rotation):
        return rot
    rot = np.array(rot)
    if rot.shape == (4,):
        return st.Rotation.from_quat(rot)
    if rot.shape == (3, 3):
        return st.Rotation.from_matrix(rot)
    if rot.shape == (3,):
        return st.Rotation.from_rotvec(rot)
    raise ValueError(msg)


# ---------------------------------------------------------------------------
# energy / time metrics
# ---------------------------------------------------------------------------

class Energy(Metric):
    """Identity on ekin (reference geom.py:300-307)."""

    def __init__(self):
        super().__init__([0], ["ekin"], ["MeV"], "MeV")


class Lethargy(Metric):
    """u = log(E0 / ekin) (reference geom.py:310-353)."""

    _param_names = ("E0",)
    _xml_dim = 1

    def __init__(self, E0=10):
        super().__init__([0], ["u"], ["[let]"], "MeV")
        self.E0 = E0

    def transform(self, ekins):
        return np.log(self.E0 / ekins)

    def inverse_transform(self, us):
        return self.E0 * np.exp(-us)

    def jac(self, ekins):
        return 1 / ekins.reshape(-1)


class Wavelength(Metric):
    """lambda[AA] = 9.045 / sqrt(E[meV]) (reference geom.py:356-380)."""

    def __init__(self):
        super().__init__([0], ["l"], ["[AA]"], "AA")

    def transform(self, ekins):
        return 9.045 / np.sqrt(ekins * 1e9)

    def inverse_transform(self, ls):
        return 81.82e-9 / ls ** 2

    def jac(self, ekins):
        # (the reference omits the reshape, geom.py:375-376, which breaks Geometry.jac)
        return 9.045 / (2 * np.sqrt(ekins.reshape(-1) ** 3))


class Time(Metric):
    """Identity on t (reference geom.py:383-390)."""

    def __init__(self):
        super().__init__([7], ["t"], ["ms"], "ms")


class Decade(Metric):
    """d = log10(t) (reference geom.py:393-418)."""

    def __init__(self):
        super().__init__([7], ["d"], ["[dec]"], "ms")

    def transform(self, ts):
        return np.log10(ts)

    def inverse_transform(self, ds):
        return 10 ** ds

    def jac(self, ts):
        return np.log10(np.e) / ts.reshape(-1)


# ---------------------------------------------------------------------------
# position metrics
# ---------------------------------------------------------------------------

class Vol(Metric):
    """3-D position, box limits (reference geom.py:421-468)."""

    _param_names = ("xmin", "xmax", "ymin", "ymax", "zmin", "zmax")
    _xml_dim = 3

    def __init__(self, xmin=-np.inf, xmax=np.inf, ymin=-np.inf, ymax=np.inf,
                 zmin=-np.inf, zmax=np.inf):
        super().__init__([1, 2, 3], ["x", "y", "z"], ["cm", "cm", "cm"], "cm^3")
        self.xmin, self.xmax = xmin, xmax
        self.ymin, self.ymax = ymin, ymax
        self.zmin, self.zmax = zmin, zmax


class _FlatXY(Metric):
    """Shared part of the flat-surface metrics that keep (x, y) and pin z."""

    _xml_dim = 2

    def transform(self, poss):
        return poss[:, :2]

    def inverse_transform(self, poss):
        z_col = np.broadcast_to(self.z, (*poss.shape[:-1], 1))
        return np.concatenate((poss, z_col), axis=1)


class SurfXY(_FlatXY):
    """Flat source in the XY plane (reference geom.py:471-520)."""

    _param_names = ("xmin", "xmax", "ymin", "ymax", "z")

    def __init__(self, xmin=-np.inf, xmax=np.inf, ymin=-np.inf, ymax=np.inf, z=0):
        super().__init__([1, 2, 3], ["x", "y"], ["cm", "cm"], "cm^2")
        self.xmin, self.xmax, self.ymin, self.ymax, self.z = xmin, xmax, ymin, ymax, z


class _Polar2D(Metric):
    _param_names = ("rho_min", "rho_max", "psi_min", "psi_max", "z")
    _xml_dim = 2

    def _set(self, rho_min, rho_max, psi_min, psi_max, z):
        self.rho_min, self.rho_max = rho_min, rho_max
        self.psi_min, self.psi_max = psi_min, psi_max
        self.z = z


class SurfR(_Polar2D):
    """(rho, psi[deg]) parametrisation (reference geom.py:523-590)."""

    def __init__(self, rho_min=0, rho_max=np.inf, psi_min=-np.pi, psi_max=np.pi, z=0):
        super().__init__([1, 2, 3], ["rho", "psi"], ["cm", "deg"], "cm")
        self._set(rho_min, rho_max, psi_min, psi_max, z)

    def transform(self, poss):
        rhos = np.sqrt(poss[:, 0] ** 2 + poss[:, 1] ** 2)
        psis = np.arctan2(poss[:, 1], poss[:, 0]) * 180 / np.pi
        return np.stack((rhos, psis), axis=1)

    def inverse_transform(self, poss):
        # (the reference stacks a (N,1) z column with (N,) x/y and raises,
        # geom.py:557-562; here all three columns are 1-D)
        z_col = np.full(len(poss), float(self.z))
        x_col = poss[:, 0] * np.cos(poss[:, 1] * np.pi / 180)
        y_col = poss[:, 0] * np.sin(poss[:, 1] * np.pi / 180)
        return np.stack((x_col, y_col, z_col), axis=1)

    def jac(self, poss):
        return 1 / np.sqrt(poss[:, 0] ** 2 + poss[:, 1] ** 2).reshape(-1)


class SurfR2(_Polar2D):
    """(rho^2, psi[deg]) parametrisation (reference geom.py:593-660)."""

    def __init__(self, rho_min=0, rho_max=np.inf, psi_min=-np.pi, psi_max=np.pi, z=0):
        super().__init__([1, 2, 3], ["rho^2", "psi"], ["cm^2", "deg"], "cm^2")
        self._set(rho_min, rho_max, psi_min, psi_max, z)

    def transform(self, poss):
        rhos = poss[:, 0] ** 2 + poss[:, 1] ** 2
        psis = np.arctan2(poss[:, 1], poss[:, 0]) * 180 / np.pi
        return np.stack((rhos, psis), axis=1)

    def inverse_transform(self, poss):
        z_col = np.full(len(poss), float(self.z))
        x_col = poss[:, 0] ** 0.5 * np.cos(poss[:, 1] * np.pi / 180)
        y_col = poss[:, 0] ** 0.5 * np.sin(poss[:, 1] * np.pi / 180)
        return np.stack((x_col, y_col, z_col), axis=1)

    def jac(self, poss):
        return np.full(len(poss[:, 0]), 2)


class SurfCircle(_Polar2D, _FlatXY):
    """(x, y) with circular limits (reference geom.py:663-715)."""

    def __init__(self, rho_min=0, rho_max=np.inf, psi_min=-np.pi, psi_max=np.pi, z=0):
        Metric.__init__(self, [1, 2, 3], ["x", "y"], ["cm", "cm"], "cm^2")
        self._set(rho_min, rho_max, psi_min, psi_max, z)


class Guide(Metric):
    """Leaks through the mirrors of a (curved) rectangular guide
    (reference geom.py:718-865): (z, t, mu, phi[deg])."""

    _param_names = ("xwidth", "yheight", "zmax", "rcurv")
    # the reference writes dim=4 but its load() checks dim != 6 (geom.py:860-864) and
    # so never loads a Guide; here the written value is accepted
    _xml_dim = 4

    @classmethod
    def load(cls, mtree):
        toks = mtree[1].text.split()
        if int(mtree[0].text) != 4 or len(toks) != 4 or int(mtree[1].attrib["nps"]) != 4:
            raise Exception("Invalid metric tree.")
        # save() writes rcurv=None literally as "None" (geom.py:849-856)
        vals = [None if t == "None" else float(t) for t in toks]
        return cls(*vals)

    def __init__(self, xwidth, yheight, zmax=np.inf, rcurv=None):
        super().__init__([1, 2, 3, 4, 5, 6], ["z", "t", "theta", "phi"],
                         ["cm", "cm", "deg", "deg"], "cm^2 sr")
        self.xwidth, self.yheight, self.zmax, self.rcurv = xwidth, yheight, zmax, rcurv

    def _mirror_masks(self, xs, ys):
        a, b = ys / self.yheight, xs / self.xwidth
        return ((a > -b) & (a < b), (a > b) & (a > -b), (a < -b) & (a > b), (a < b) & (a < -b))

    def transform(self, posdirs):
        xs, ys, zs, dxs, dys, dzs = posdirs.T
        if self.rcurv is not None:
            rs = np.sqrt((self.rcurv + xs) ** 2 + zs ** 2)
            xs = np.sign(self.rcurv) * rs - self.rcurv
            zs = np.abs(self.rcurv) * np.arcsin(zs / rs)
            angs = zs / self.rcurv
            dxs, dzs = (dxs * np.cos(angs) + dzs * np.sin(angs),
                        -dxs * np.sin(angs) + dzs * np.cos(angs))
        masks = self._mirror_masks(xs, ys)
        yh, xw = self.yheight, self.xwidth
        t_of = (lambda m: 0.5 * yh + ys[m], lambda m: yh + 0.5 * xw - xs[m],
                lambda m: 1.5 * yh + xw - ys[m], lambda m: 2.0 * yh + 1.5 * xw + xs[m])
        mu_of = (lambda m: dxs[m], lambda m: dys[m], lambda m: -dxs[m], lambda m: -dys[m])
        ph_of = (lambda m: np.arctan2(-dys[m], dzs[m]), lambda m: np.arctan2(dxs[m], dzs[m]),
                 lambda m: np.arctan2(dys[m], dzs[m]), lambda m: np.arctan2(-dxs[m], dzs[m]))
        ts, mus, phis = np.empty_like(xs), np.empty_like(dxs), np.empty_like(dxs)
        for k, m in enumerate(masks):
            ts[m], mus[m], phis[m] = t_of[k](m), mu_of[k](m), ph_of[k](m)
        phis *= _DEG
        return np.stack((zs, ts, mus, phis), axis=1)

    def inverse_transform(self, posdirs):
        zs, ts, mus, phis = posdirs.T
        phis *= np.pi / 180  # in place, as the reference (geom.py:795)
        yh, xw = self.yheight, self.xwidth
        masks = (ts < yh, (ts > yh) & (ts < yh + xw),
                 (ts > yh + xw) & (ts < 2 * yh + xw), ts > 2 * yh + xw)
        xs, ys = np.empty_like(ts), np.empty_like(ts)
        dxs, dys, dzs = np.empty_like(mus), np.empty_like(mus), np.empty_like(mus)
        sq = np.sqrt(1 - mus ** 2)
        m = masks[0]
        xs[m], ys[m] = xw / 2, ts[m] - 0.5 * yh
        dxs[m], dzs[m], dys[m] = mus[m], sq[m] * np.cos(phis[m]), -sq[m] * np.sin(phis[m])
        m = masks[1]
        ys[m], xs[m] = yh / 2, -ts[m] + yh + 0.5 * xw
        dys[m], dzs[m], dxs[m] = mus[m], sq[m] * np.cos(phis[m]), sq[m] * np.sin(phis[m])
        m = masks[2]
        xs[m], ys[m] = -xw / 2, -ts[m] + 1.5 * yh + xw
        dxs[m], dzs[m], dys[m] = -mus[m], sq[m] * np.cos(phis[m]), sq[m] * np.sin(phis[m])
        m = masks[3]
        ys[m], xs[m] = -yh / 2, ts[m] - 2.0 * yh - 1.5 * xw
        dys[m], dzs[m], dxs[m] = -mus[m], sq[m] * np.cos(phis[m]), -sq[m] * np.sin(phis[m])
        if self.rcurv is not None:
            rs = (self.rcurv + xs) * np.sign(self.rcurv)
            angs = zs / self.rcurv
            xs = np.sign(self.rcurv) * rs * np.cos(angs) - self.rcurv
            zs = rs * np.sin(np.abs(angs))
            dxs, dzs = (dxs * np.cos(angs) - dzs * np.sin(angs),
                        dxs * np.sin(angs) + dzs * np.cos(angs))
        return np.stack((xs, ys, zs, dxs, dys, dzs), axis=1)


# ---------------------------------------------------------------------------
# direction metrics
# ---------------------------------------------------------------------------

class Isotrop(Metric):
    """Identity on the unit direction; pooled std, normalised mean
    (reference geom.py:868-966)."""

    _param_names = ("keep_xdir", "keep_ydir", "keep_zdir")
    _int_params = True
    _xml_dim = 3

    def __init__(self, keep_xdir=False, keep_ydir=False, keep_zdir=False):
        super().__init__([4, 5, 6], ["dx", "dy", "dz"], ["[dir]", "[dir]", "[dir]"], "[dir]^3")
        self.keep_xdir, self.keep_ydir, self.keep_zdir = keep_xdir, keep_ydir, keep_zdir

    def mean(self, dirs=None, vecs=None, weights=None):
        if vecs is None:
            vecs = self.transform(dirs)
        mn = np.average(vecs, axis=0, weights=weights)
        norm = np.linalg.norm(mn)
        return mn / (norm if norm != 0 else 1)

    def std(self, dirs=None, vecs=None, weights=None):
        if vecs is None:
            vecs = self.transform(dirs)
        mn = self.mean(vecs=vecs, weights=weights)
        pooled = np.sqrt(np.mean(np.average((vecs - mn) ** 2, axis=0, weights=weights)))
        return np.array(3 * [pooled])

    def mean_sharded(self, vecs, weights, allsum):
        mn = self._wavg_sharded(vecs, weights, allsum)
        norm = np.linalg.norm(mn)
        return mn / (norm if norm != 0 else 1)

    def std_from_moments(self, avg):
        mn = avg(None, 1)
        norm = np.linalg.norm(mn)
        mn = mn / (norm if norm != 0 else 1)
        return np.array(3 * [np.sqrt(np.mean(avg(mn, 2)))])

    def std_sharded(self, vecs, weights, allsum):
        mn = self.mean_sharded(vecs, weights, allsum)
        pooled = np.sqrt(np.mean(self._wavg_sharded((vecs - mn) ** 2, weights, allsum)))
        return np.array(3 * [pooled])


class Polar(Metric):
    """(theta, phi) in degrees (reference geom.py:969-1004)."""

    def __init__(self):
        super().__init__([4, 5, 6], ["theta", "phi"], ["deg", "deg"], "sr^2")

    def transform(self, dirs):
        return np.stack((np.arccos(dirs[:, 2]) * 180 / np.pi,
                         np.arctan2(dirs[:, 1], dirs[:, 0]) * 180 / np.pi), axis=1)

    def inverse_transform(self, tps):
        thetas, phis = tps.T[0] * np.pi / 180, tps.T[1] * np.pi / 180
        return np.stack((np.sin(thetas) * np.cos(phis), np.sin(thetas) * np.sin(phis),
                         np.cos(thetas)), axis=1)

    def jac(self, dirs):
        return (180 / np.pi) ** 2 / np.sin(np.arccos(dirs[:, 2]))


class PolarMu(Metric):
    """(mu = cos theta, phi[deg]) (reference geom.py:1007-1039)."""

    def __init__(self):
        super().__init__([4, 5, 6], ["mu", "phi"], ["", "deg"], "sr")

    def transform(self, dirs):
        return np.stack((dirs[:, 2], np.arctan2(dirs[:, 1], dirs[:, 0]) * 180 / np.pi), axis=1)

    def inverse_transform(self, tps):
        mus, phis = tps.T
        sq = np.sqrt(1 - mus ** 2)
        return np.stack((sq * np.cos(phis * np.pi / 180), sq * np.sin(phis * np.pi / 180), mus),
                        axis=1)


_metrics = {name: globals()[name] for name in METRIC_ORDER}


# ---------------------------------------------------------------------------
# usual geometries (reference geom.py:1062-1163)
# ---------------------------------------------------------------------------

def GeomFlat(xmin=-np.inf, xmax=np.inf, ymin=-np.inf, ymax=np.inf, z=0, E0=10,
             keep_zdir=True, trasl=None, rot=None):
    """Lethargy + flat XY surface + isotropic direction."""
    return Geometry([Lethargy(E0), SurfXY(xmin, xmax, ymin, ymax, z),
                     Isotrop(keep_zdir=keep_zdir)], trasl=trasl, rot=rot)


def GeomFlatTemp(xmin=-np.inf, xmax=np.inf, ymin=-np.inf, ymax=np.inf, z=0, E0=10,
                 keep_zdir=True, trasl=None, rot=None):
    """:func:`GeomFlat` plus the decade time metric."""
    return Geometry([Lethargy(E0), SurfXY(xmin, xmax, ymin, ymax, z),
                     Isotrop(keep_zdir=keep_zdir), Decade()], trasl=trasl, rot=rot)


def GeomGuide(xwidth, yheight, zmax=np.inf, rcurv=None, E0=10, trasl=None, rot=None):
    """Lethargy + neutron-guide mirrors."""
    return Geometry([Lethargy(E0), Guide(xwidth, yheight, zmax, rcurv)], trasl=trasl, rot=rot)


def GeomActiv(xmin=-np.inf, xmax=np.inf, ymin=-np.inf, ymax=np.inf, zmin=-np.inf,
              zmax=np.inf, trasl=None, rot=None):
    """Energy + volume + isotropic direction (activation sources)."""
    return Geometry([Energy(), Vol(xmin, xmax, ymin, ymax, zmin, zmax), Isotrop()],
                    trasl=trasl, rot=rot)
