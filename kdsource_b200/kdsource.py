"""KDSource object: KDE particle source on an MCPL list, GPU evaluated.

API-compatible with the hot-path part of the reference's
``python/kdsource/kdsource.py`` (:32-321): ``KDSource(plist, geom, bw, J,
kernel)``, ``fit``, ``evaluate``, ``save`` and the module function ``load``,
with the same XML source-file format.  Kernel sums run in the CUDA kernels
behind :class:`kdsource_b200.treekde.TreeKDE`; bandwidth selection in
:mod:`kdsource_b200.kde`.  The matplotlib ``plot_*`` helpers of the reference
are presentation code and are not part of this package.
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

    def fit(self, N=-1, skip=0, scaling=None, importance=None, bw_method=None, **kwargs):
        """Load up to N particles (after `skip`) and fit the KDE model
        (reference kdsource.py:132-213): scaling = weighted std of each KDE
        variable / importance (or the given scaling; zeros -> 1), then the
        bandwidth method, if any, on the scaled variables."""
        parts, ws = self.plist.get(N, skip)
        N = len(parts)
        if N == 0:
            raise Exception("No particles for fitting.")
        print("Using {} particles for fit.".format(N))
        vecs = self.geom.transform(parts)
        self.N_eff = np.sum(ws) ** 2 / np.sum(ws ** 2)
        if scaling is None:
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
        dens = np.zeros(M)
        jacs = np.zeros(M)
        for s in range(lo, hi, chunk):
            n = min(chunk, hi - s)
            d_parts = _lib.to_dev(parts[s:s + n])
            d_vecs = _lib.empty(n * D, t.float64)
            d_jac = _lib.empty(n, t.float64)
            _lib.check(L.kdsb200_geom_transform(_lib.ptr(d_parts), n, ctypes.byref(gs), rot_p,
                                                inv_p, _lib.ptr(d_vecs), _lib.ptr(d_jac),
                                                _lib.stream()))
            d_out = self.kde.evaluate_device(d_vecs, n, dev)
            dens[s:s + n] = d_out.cpu().numpy()
            jacs[s:s + n] = d_jac.cpu().numpy()
        if world > 1:
            dens = _dist.allreduce_sum_numpy(dens)
            jacs = _dist.allreduce_sum_numpy(jacs)
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
                bw = bw * factor
                self.kde.bw = bw
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
