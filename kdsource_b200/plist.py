"""Particle lists: MCPL files wrapped for KDSource.

API-compatible with ``PList`` of the reference's ``python/kdsource/plist.py``
(:240-429): same constructor arguments, ``set_params``, ``get``, ``save``,
``load``.  MCPL files are read with this package's own MCPL-3 reader
(:mod:`kdsource_b200.mcpl`); the format converters that shell out to the
legacy ``kdtool`` (ssw/phits/ptrac/stock/ssv -> mcpl, plist.py:17-237) are not
part of the kernel-density path and only pass MCPL input through.
"""

import os
from xml.etree.ElementTree import SubElement

import numpy as np

from . import mcpl
from .geom import _as_rotation
from .utils import pt2pdg


def convert2mcpl(filename, readformat):
    """Return the MCPL file for `filename` (reference plist.py:17-75).

    Only ``readformat="mcpl"`` (or an already converted sibling ``.mcpl[.gz]``)
    is supported; other formats need the legacy ``kdtool`` converters."""
    if readformat == "mcpl":
        print("Using existing file {}".format(filename))
        return filename
    stem = filename.split(".")[0]
    for ext in (".mcpl.gz", ".mcpl"):
        if os.path.isfile(stem + ext):
            print("Using existing file {}".format(stem + ext))
            return stem + ext
    if readformat not in ["ssw", "phits", "ptrac", "stock", "ssv"]:
        raise Exception("Invalid {} format".format(readformat))
    raise NotImplementedError(
        "conversion from {} needs the legacy kdtool, which this package does not ship; "
        "convert to MCPL first".format(readformat))


def join2mcpl(filenames, readformat):
    """Single-file pass-through of the reference's join2mcpl (plist.py:78-125)."""
    if np.isscalar(filenames):
        filenames = [filenames]
    names = [convert2mcpl(n, readformat) for n in filenames]
    if len(names) == 1:
        return names[0]
    raise NotImplementedError("joining several MCPL files needs mcpltool (not shipped)")


class PList:
    def __init__(self, filename, readformat="mcpl", pt="n", trasl=None, rot=None,
                 switch_x2z=False, set_params=True):
        """MCPL particle list with an optional rigid transformation.

        ``trasl`` is added to positions, then ``rot`` (rotation vector, quaternion,
        matrix or scipy Rotation) is applied to positions and directions, then
        ``switch_x2z`` permutes (x,y,z) -> (y,z,x)."""
        if np.isscalar(filename):
            self.filename = convert2mcpl(filename, readformat)
        else:
            self.filename = join2mcpl(filename, readformat)
        if trasl is not None:
            trasl = np.array(trasl).reshape(-1)
            if trasl.shape != (3,):
                raise ValueError("Invalid trasl")
        self.pt = pt
        self.trasl = trasl
        self.rot = _as_rotation(rot, "Invalid rot")
        self.x2z = switch_x2z
        self.params_set = False
        if set_params:
            self.set_params()

    def _valid_mask(self, block):
        return np.logical_and(block.weight > 0, block.pdgcode == pt2pdg(self.pt))

    def set_params(self):
        """Totals over the valid particles of the whole file: sum of weights,
        sum of squared weights, N and N_eff (reference plist.py:318-347)."""
        sum_weights = p2 = N = 0
        with mcpl.MCPLFile(self.filename, blocklength=1 << 20) as pl:
            for pb in pl.particle_blocks:
                ws = pb.weight[self._valid_mask(pb)]
                N += len(ws)
                sum_weights += ws.sum()
                p2 += (ws ** 2).sum()
        if N == 0:
            raise Exception("Empty particle list")
        N_eff = sum_weights ** 2 / p2
        print("sum_weights = {}\np2 = {}\nN = {}\nN_eff = {}".format(
            sum_weights, p2, N, N_eff))
        self.sum_weights = sum_weights
        self.p2 = p2
        self.N = N
        self.N_eff = N_eff
        self.params_set = True

    def get(self, N=-1, skip=0):
        """Up to N particles after skipping `skip`: ``(parts (n,8), ws (n,))`` with
        columns [ekin,x,y,z,dx,dy,dz,t]; invalid particles (w<=0 or other pdg
        code) are dropped (reference plist.py:349-395)."""
        if self._use_device_decoder(N, skip):
            # large lists: records decoded on the GPU (bit-identical, see get_device)
            parts, ws = self.get_device(N, skip)
            return parts.cpu().numpy(), ws.cpu().numpy()
        with mcpl.MCPLFile(self.filename) as pl:
            if N < 0:
                N = pl.nparticles
            pl.blocklength = N
            pl.skip_forward(skip)
            pb = pl.read_block()
        if pb is None:
            # the reference returns these in swapped order (plist.py:378-379);
            # here the order matches the non-empty case
            return np.zeros((0, 8)), np.zeros(0)
        parts = np.stack((pb.ekin, pb.x, pb.y, pb.z, pb.ux, pb.uy, pb.uz, pb.time), axis=1)
        keep = self._valid_mask(pb)
        parts = parts[keep]
        ws = pb.weight[keep]
        if self.trasl is not None:
            parts[:, 1:4] += self.trasl
        if self.rot is not None:
            parts[:, 1:4] = self.rot.apply(parts[:, 1:4])
            parts[:, 4:7] = self.rot.apply(parts[:, 4:7])
        if self.x2z:
            parts = parts[:, [0, 2, 3, 1, 5, 6, 4, 7]]
        return parts, ws

    DEVICE_DECODE_MIN = 1 << 20  # records; below this the host decoder is as fast

    def _use_device_decoder(self, N, skip):
        try:
            with mcpl.MCPLFile(self.filename) as pl:
                n = max(0, pl.nparticles - skip)
                little = pl.hdr.endian == "<"
            n = n if N < 0 else min(N, n)
            if n < self.DEVICE_DECODE_MIN or not little:
                return False
            import torch

            return torch.cuda.is_available()
        except Exception:
            return False

    def get_device(self, N=-1, skip=0):
        """:meth:`get` for device-side consumers (the sampler): the records are uploaded
        as they lie in the file and decoded by ``kdsb200_mcpl_unpack``; returns CUDA
        tensors ``(parts (n,8) float64, ws (n,) float64)`` with the same values as
        :meth:`get`.  Filtering and the rigid transformation use torch tensor ops
        (buffer plumbing, not a compute path)."""
        import ctypes

        from . import _lib

        L = _lib.load()
        t = _lib.torch()
        with mcpl.MCPLFile(self.filename) as pl:
            lay = pl.device_layout()
            if lay is None:  # big-endian file: host decoder
                parts, ws = self.get(N, skip)
                return _lib.to_dev(parts), _lib.to_dev(ws)
            if N < 0:
                N = pl.nparticles
            pl.skip_forward(skip)
            buf, n = pl.read_raw_bytes(N)
        if n == 0:
            return _lib.empty(0, t.float64).reshape(0, 8), _lib.empty(0, t.float64)
        d_raw = t.from_numpy(buf).cuda()
        parts = _lib.empty(n * 8, t.float64)
        ws = _lib.empty(n, t.float64)
        pdg = _lib.empty(n, t.int32)
        _lib.check(L.kdsb200_mcpl_unpack(_lib.ptr(d_raw), n, ctypes.byref(lay), _lib.ptr(parts),
                                         _lib.ptr(ws), _lib.ptr(pdg), _lib.stream()))
        keep = (ws > 0) & (pdg == pt2pdg(self.pt))
        parts = parts.reshape(n, 8)[keep]
        ws = ws[keep]
        if self.trasl is not None:
            parts[:, 1:4] += t.as_tensor(self.trasl, dtype=t.float64, device="cuda")
        if self.rot is not None:
            R = t.as_tensor(self.rot.as_matrix(), dtype=t.float64, device="cuda")
            parts[:, 1:4] = parts[:, 1:4] @ R.T
            parts[:, 4:7] = parts[:, 4:7] @ R.T
        if self.x2z:
            parts = parts[:, [0, 2, 3, 1, 5, 6, 4, 7]]
        return parts.contiguous(), ws.contiguous()

    def rotvec(self):
        return None if self.rot is None else np.asarray(self.rot.as_rotvec(), dtype=np.float64)

    def save(self, pltree):
        """Write the <PList> children (reference plist.py:397-411)."""
        SubElement(pltree, "pt").text = self.pt
        SubElement(pltree, "mcplname").text = os.path.abspath(self.filename)
        SubElement(pltree, "trasl").text = (
            np.array_str(self.trasl)[1:-1] if self.trasl is not None else "")
        SubElement(pltree, "rot").text = (
            np.array_str(self.rot.as_rotvec())[1:-1] if self.rot is not None else "")
        SubElement(pltree, "x2z").text = str(int(self.x2z))

    @staticmethod
    def load(pltree, set_params=True):
        """Build a PList from a <PList> element (reference plist.py:413-429).
        ``set_params=False`` skips the pass over the whole file that computes
        N / N_eff (the sampler does not need them)."""
        def vec(el):
            return np.array(el.text.split(), dtype="float64") if el.text else None

        return PList(pltree[1].text, pt=pltree[0].text, trasl=vec(pltree[2]),
                     rot=vec(pltree[3]), switch_x2z=bool(int(pltree[4].text)),
                     set_params=set_params)


def savessv(*args, **kwargs):
    raise NotImplementedError("SSV text export is outside the kernel-density path")


def appendssv(*args, **kwargs):
    raise NotImplementedError("SSV text export is outside the kernel-density path")
