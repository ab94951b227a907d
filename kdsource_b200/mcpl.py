"""MCPL-3 particle-list I/O (host side, numpy).

The reference reads particle lists through the third-party ``mcpl`` Python
module (``python/kdsource/plist.py:292-301,349-360``) and writes them through
libmcpl (``clib/src/writemcpl.c:15-69``, ``clib/src/resamplemcpl.c:30-44``).
Neither library ships with the reference tree, so this module implements the
MCPL-3 on-disk format directly.  ``MCPLFile`` mirrors the subset of the
``mcpl.MCPLFile`` interface that ``plist.py`` uses: ``nparticles``,
``particle_blocks``, ``read_block()``, ``skip_forward()`` and block attributes
``ekin,x,y,z,ux,uy,uz,time,weight,pdgcode`` (+ ``polx/poly/polz/userflags``).

On-disk layout (little or big endian, flagged in the magic):

    "MCPL" "003" 'L'|'B'
    u64 nparticles
    u32 ncomments, nblobs, opt_userflags, opt_polarisation, opt_singleprec
    i32 opt_universalpdgcode
    u32 particlesize
    u32 opt_universalweight  (+ f64 value when set)
    string srcname ; comments ; blob keys ; blobs   (string = u32 len + bytes)
    particle records: [pol x3 FP]? pos x3 FP, dirpack x2 FP, ekin FP (sign bit
                      carries one direction bit), time FP, [weight FP]?,
                      [pdgcode i32]?, [userflags u32]?

Directions use MCPL's "adaptive projection packing": the two stored numbers are
(x, y), (1/z, y) or (x, 1/z) depending on which component is largest, and the
sign of the dropped component rides on the sign bit of ``ekin``.
"""

import gzip
import os
import struct

import numpy as np

__all__ = [
    "MCPLFile",
    "MCPLHeader",
    "ParticleBlock",
    "pack_direction",
    "unpack_direction",
    "record_dtype",
    "build_header",
    "write_mcpl_file",
    "pack_records",
    "patch_nparticles",
]


class MCPLHeader:
    """Decoded MCPL header (format version 2 and 3 share this layout)."""

    def __init__(self):
        self.version = 3
        self.endian = "<"
        self.nparticles = 0
        self.opt_userflags = False
        self.opt_polarisation = False
        self.opt_singleprec = True
        self.opt_universalpdgcode = 0
        self.opt_universalweight = 0.0
        self.particlesize = 0
        self.srcname = ""
        self.comments = []
        self.blobs = {}
        self.headersize = 0


def record_dtype(
    singleprec=True,
    polarisation=False,
    universal_pdg=False,
    universal_weight=False,
    userflags=False,
    endian="<",
):
    """numpy structured dtype of one particle record for the given options."""
    fp = endian + ("f4" if singleprec else "f8")
    fields = []
    if polarisation:
        fields.append(("pol", fp, (3,)))
    fields.append(("pos", fp, (3,)))
    fields.append(("dirpack", fp, (2,)))
    fields.append(("ekin", fp))
    fields.append(("time", fp))
    if not universal_weight:
        fields.append(("weight", fp))
    if not universal_pdg:
        fields.append(("pdgcode", endian + "i4"))
    if userflags:
        fields.append(("userflags", endian + "u4"))
    return np.dtype(fields)


def pack_direction(ux, uy, uz, ekin):
    """Adaptive projection packing of unit vectors.

    Returns (p0, p1, signed_ekin) as float64 arrays.  Mirrors libmcpl's
    ``mcpl_unitvect_pack_adaptproj`` as used by ``mcpl_add_particle`` (called
    from ``clib/src/resamplemcpl.c:42`` and ``clib/src/writemcpl.c:66``).
    """
    ux = np.asarray(ux, dtype=np.float64)
    uy = np.asarray(uy, dtype=np.float64)
    uz = np.asarray(uz, dtype=np.float64)
    ekin = np.abs(np.asarray(ekin, dtype=np.float64))
    ax, ay, az = np.abs(ux), np.abs(uy), np.abs(uz)
    zsmall = az < np.maximum(ax, ay)
    with np.errstate(divide="ignore"):
        invz = np.where(uz != 0.0, 1.0 / np.where(uz != 0.0, uz, 1.0), np.inf)
    xbig = zsmall & (ax >= ay)
    ybig = zsmall & ~(ax >= ay)
    p0 = np.where(xbig, invz, ux)
    p1 = np.where(ybig, invz, uy)
    signsrc = np.where(xbig, ux, np.where(ybig, uy, uz))
    sek = np.copysign(ekin, signsrc)
    return p0, p1, sek


def unpack_direction(p0, p1, signed_ekin):
    """Inverse of :func:`pack_direction`; returns (ux, uy, uz, ekin)."""
    p0 = np.asarray(p0, dtype=np.float64)
    p1 = np.asarray(p1, dtype=np.float64)
    sek = np.asarray(signed_ekin, dtype=np.float64)
    sgn = np.where(np.signbit(sek), -1.0, 1.0)
    c0 = np.abs(p0) > 1.0
    c1 = ~c0 & (np.abs(p1) > 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv0 = 1.0 / p0
        inv1 = 1.0 / p1
    # case 0: (1/z, y) + sign(x)
    z0 = inv0
    x0 = sgn * np.sqrt(np.maximum(0.0, 1.0 - (p1 * p1 + z0 * z0)))
    # case 1: (x, 1/z) + sign(y)
    z1 = inv1
    y1 = sgn * np.sqrt(np.maximum(0.0, 1.0 - (p0 * p0 + z1 * z1)))
    # case 2: (x, y) + sign(z)
    z2 = sgn * np.sqrt(np.maximum(0.0, 1.0 - (p0 * p0 + p1 * p1)))
    ux = np.where(c0, x0, p0)
    uy = np.where(c1, y1, p1)
    uz = np.where(c0, z0, np.where(c1, z1, z2))
    return ux, uy, uz, np.abs(sek)


class ParticleBlock:
    """A block of decoded particles (float64 columns, like ``mcpl`` blocks)."""

    def __init__(self, rec, hdr):
        n = len(rec)
        pos = rec["pos"].astype(np.float64)
        self.x = np.ascontiguousarray(pos[:, 0])
        self.y = np.ascontiguousarray(pos[:, 1])
        self.z = np.ascontiguousarray(pos[:, 2])
        dp = rec["dirpack"].astype(np.float64)
        ux, uy, uz, ek = unpack_direction(
            dp[:, 0], dp[:, 1], rec["ekin"].astype(np.float64)
        )
        self.ux, self.uy, self.uz, self.ekin = ux, uy, uz, ek
        self.time = rec["time"].astype(np.float64)
        if hdr.opt_universalweight:
            self.weight = np.full(n, hdr.opt_universalweight, dtype=np.float64)
        else:
            self.weight = rec["weight"].astype(np.float64)
        if hdr.opt_universalpdgcode:
            self.pdgcode = np.full(n, hdr.opt_universalpdgcode, dtype=np.int32)
        else:
            self.pdgcode = rec["pdgcode"].astype(np.int32)
        if hdr.opt_polarisation:
            pol = rec["pol"].astype(np.float64)
            self.polx, self.poly, self.polz = pol[:, 0], pol[:, 1], pol[:, 2]
        else:
            self.polx = self.poly = self.polz = np.zeros(n)
        if hdr.opt_userflags:
            self.userflags = rec["userflags"].astype(np.uint32)
        else:
            self.userflags = np.zeros(n, dtype=np.uint32)

    def __len__(self):
        return len(self.x)


# ---- indexed gzip members ----------------------------------------------------
# The files this package writes (device gzip, csrc/gzip.cu; host writer below) are
# sequences of gzip members that each announce their own compressed size in an FEXTRA
# subfield "KD" (4 bytes, little endian) -- the idea of BGZF's "BC" field with room for
# members beyond 64 KB.  A reader hops from header to header, takes every member's raw
# size from its ISIZE trailer, and inflates the members in parallel (zlib releases the
# GIL).  Any other gzip file goes through Python's gzip module as before.
_KD_HEAD = struct.Struct("<4sIBBH2sHI")  # magic+CM+FLG, MTIME, XFL, OS, XLEN, "KD", LEN, size


def _kd_member_size(buf, off):
    """Compressed size of the member starting at buf[off] if it carries the "KD" subfield
    as its first (and only) extra field, else None."""
    if off + 20 > len(buf):
        return None
    magic, _, _, _, xlen, si, ln, size = _KD_HEAD.unpack_from(buf, off)
    if magic[:3] != b"\x1f\x8b\x08" or not (magic[3] & 4) or xlen != 8 or si != b"KD" or ln != 4:
        return None
    if size < 28 or off + size > len(buf):
        return None
    return size


def gz_member_index(buf):
    """(offsets, sizes, raw sizes) of the members of an indexed gzip file held in `buf`
    (bytes-like / mmap), or None if any member lacks the "KD" size field."""
    offs, sizes, raws = [], [], []
    off, n = 0, len(buf)
    while off < n:
        size = _kd_member_size(buf, off)
        if size is None:
            return None
        offs.append(off)
        sizes.append(size)
        raws.append(struct.unpack_from("<I", buf, off + size - 4)[0])
        off += size
    if not offs:
        return None
    return (np.asarray(offs, dtype=np.int64), np.asarray(sizes, dtype=np.int64),
            np.asarray(raws, dtype=np.int64))


def _inflate_workers():
    return max(1, min(32, (os.cpu_count() or 1)))


class _IndexedGzReader:
    """File-like reader (read / readinto / close) over an indexed multi-member gzip file:
    members are inflated a window at a time on a thread pool."""

    WINDOW_BYTES = 64 << 20  # raw bytes inflated per window

    def __init__(self, filename, mm, index):
        from concurrent.futures import ThreadPoolExecutor

        self._fh = open(filename, "rb")
        self._mm = mm
        self._offs, self._sizes, self._raws = index
        self._next = 0          # next member to inflate
        self._cum = None        # cumulative raw sizes (built on the first skip)
        self._buf = b""         # inflated, not yet consumed
        self._bpos = 0
        self._pool = ThreadPoolExecutor(max_workers=_inflate_workers())

    def _inflate(self, k):
        import zlib

        o, n = int(self._offs[k]), int(self._sizes[k])
        out = zlib.decompress(self._mm[o:o + n], 31)
        if len(out) != int(self._raws[k]) and int(self._raws[k]) < (1 << 32) - 1:
            raise IOError("gzip member {}: size mismatch".format(k))
        return out

    def _fill(self, want):
        """Make at least `want` unconsumed bytes available (or everything that is left)."""
        have = len(self._buf) - self._bpos
        if have >= want or self._next >= len(self._offs):
            return
        k0 = self._next
        k1, raw = k0, 0
        need = max(want - have, self.WINDOW_BYTES)
        while k1 < len(self._offs) and raw < need:
            raw += int(self._raws[k1])
            k1 += 1
        parts = list(self._pool.map(self._inflate, range(k0, k1), chunksize=max(1, (k1 - k0) // 64)))
        self._next = k1
        self._buf = self._buf[self._bpos:] + b"".join(parts)
        self._bpos = 0

    def read(self, n=-1):
        if n is None or n < 0:
            self._fill(1 << 62)
            n = len(self._buf) - self._bpos
        else:
            self._fill(n)
        out = self._buf[self._bpos:self._bpos + n]
        self._bpos += len(out)
        return out

    def skip(self, n):
        """Advance by n inflated bytes; members that lie wholly inside the skipped range are
        never inflated (a rank of a sharded read hops straight to its share)."""
        n = int(n)
        take = min(len(self._buf) - self._bpos, n)
        self._bpos += take
        n -= take
        if n <= 0:
            return
        self._buf, self._bpos = b"", 0  # drained: the stream now stands at member _next
        if self._cum is None:
            self._cum = np.cumsum(self._raws)
        start = int(self._cum[self._next - 1]) if self._next else 0
        k = int(np.searchsorted(self._cum, start + n, side="right"))  # members < k end inside
        if k > self._next:
            n -= int(self._cum[k - 1]) - start
            self._next = k
        if n > 0:
            self.read(n)

    def readinto(self, b):
        # window by window: a multi-gigabyte read never holds more than one window of
        # inflated bytes besides the destination
        view = memoryview(b).cast("B")
        done = 0
        while done < len(view):
            data = self.read(min(len(view) - done, self.WINDOW_BYTES))
            if not data:
                break
            view[done:done + len(data)] = data
            done += len(data)
        return done

    def close(self):
        self._pool.shutdown(wait=False)
        try:
            self._mm.close()
        finally:
            self._fh.close()


def _open_maybe_gz(filename):
    with open(filename, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b":
        import mmap

        with open(filename, "rb") as fh:
            mm = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
        index = gz_member_index(mm) if _kd_member_size(mm, 0) is not None else None
        if index is not None:
            return _IndexedGzReader(filename, mm, index)
        mm.close()
        return gzip.open(filename, "rb")
    return open(filename, "rb")


def gz_indexed_members(data, member_bytes=1 << 20, compresslevel=6):
    """`data` (bytes-like) as a sequence of indexed gzip members of `member_bytes` raw bytes,
    deflated in parallel; readable by any gzip reader, in parallel by `_IndexedGzReader`."""
    import zlib
    from concurrent.futures import ThreadPoolExecutor

    view = memoryview(data).cast("B")

    def one(lo):
        chunk = view[lo:lo + member_bytes]
        co = zlib.compressobj(compresslevel, zlib.DEFLATED, -15)
        body = co.compress(chunk) + co.flush()
        size = 20 + 1 + len(body) + 8  # header + empty FNAME's NUL + deflate + CRC32 + ISIZE
        head = _KD_HEAD.pack(b"\x1f\x8b\x08\x0c", 0, 0, 255, 8, b"KD", 4, size) + b"\x00"
        return head + body + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk) & 0xFFFFFFFF)

    starts = range(0, max(len(view), 1), member_bytes)
    with ThreadPoolExecutor(max_workers=_inflate_workers()) as pool:
        return list(pool.map(one, starts))


def _read_exact(fh, n):
    buf = fh.read(n)
    if len(buf) != n:
        raise IOError("MCPL: unexpected end of file")
    return buf


def _read_string(fh, endian):
    (n,) = struct.unpack(endian + "I", _read_exact(fh, 4))
    return _read_exact(fh, n)


class MCPLFile:
    """Sequential reader for ``.mcpl`` / ``.mcpl.gz`` files."""

    def __init__(self, filename, blocklength=10000):
        self.filename = str(filename)
        self.blocklength = int(blocklength)
        self._fh = _open_maybe_gz(self.filename)
        self.hdr = self._read_header()
        self.nparticles = self.hdr.nparticles
        self._pos = 0
        self._dtype = record_dtype(
            singleprec=self.hdr.opt_singleprec,
            polarisation=self.hdr.opt_polarisation,
            universal_pdg=bool(self.hdr.opt_universalpdgcode),
            universal_weight=bool(self.hdr.opt_universalweight),
            userflags=self.hdr.opt_userflags,
            endian=self.hdr.endian,
        )
        if self._dtype.itemsize != self.hdr.particlesize:
            raise IOError(
                "MCPL: particle size {} in header does not match options "
                "({})".format(self.hdr.particlesize, self._dtype.itemsize)
            )

    def _read_header(self):
        fh = self._fh
        h = MCPLHeader()
        magic = _read_exact(fh, 8)
        if magic[:4] != b"MCPL":
            raise IOError("Not an MCPL file: {}".format(self.filename))
        try:
            h.version = int(magic[4:7])
        except ValueError:
            raise IOError("MCPL: bad version field")
        if h.version not in (2, 3):
            raise IOError("MCPL: unsupported format version {}".format(h.version))
        if magic[7:8] == b"L":
            h.endian = "<"
        elif magic[7:8] == b"B":
            h.endian = ">"
        else:
            raise IOError("MCPL: bad endianness field")
        e = h.endian
        (h.nparticles,) = struct.unpack(e + "Q", _read_exact(fh, 8))
        arr = struct.unpack(e + "IIIIIiII", _read_exact(fh, 32))
        ncomments, nblobs = arr[0], arr[1]
        h.opt_userflags = bool(arr[2])
        h.opt_polarisation = bool(arr[3])
        h.opt_singleprec = bool(arr[4])
        h.opt_universalpdgcode = int(arr[5])
        h.particlesize = int(arr[6])
        size = 48
        if arr[7]:
            (h.opt_universalweight,) = struct.unpack(e + "d", _read_exact(fh, 8))
            size += 8
        s = _read_string(fh, e)
        size += 4 + len(s)
        h.srcname = s.decode("utf8", "replace")
        for _ in range(ncomments):
            s = _read_string(fh, e)
            size += 4 + len(s)
            h.comments.append(s.decode("utf8", "replace"))
        keys = []
        for _ in range(nblobs):
            s = _read_string(fh, e)
            size += 4 + len(s)
            keys.append(s.decode("utf8", "replace"))
        for k in keys:
            s = _read_string(fh, e)
            size += 4 + len(s)
            h.blobs[k] = s
        h.headersize = size
        return h

    # -- interface used by PList (plist.py:292-301, 349-360) ---------------
    def skip_forward(self, n):
        n = int(min(max(n, 0), self.nparticles - self._pos))
        if n:
            if hasattr(self._fh, "skip"):  # indexed gzip: whole members are hopped over
                self._fh.skip(n * self._dtype.itemsize)
            else:
                self._fh.read(n * self._dtype.itemsize)
            self._pos += n

    def read_raw(self, n):
        """Read up to n raw records as a structured array (None at EOF)."""
        n = int(min(n, self.nparticles - self._pos))
        if n <= 0:
            return None
        buf = _read_exact(self._fh, n * self._dtype.itemsize)
        self._pos += n
        return np.frombuffer(buf, dtype=self._dtype, count=n)

    def read_raw_bytes(self, n):
        """Up to n records as they lie in the file: (uint8 array, count) or (None, 0)."""
        n = int(min(n, self.nparticles - self._pos))
        if n <= 0:
            return None, 0
        buf = np.empty(n * self._dtype.itemsize, dtype=np.uint8)
        view = memoryview(buf)
        got = 0
        while got < len(buf):
            k = self._fh.readinto(view[got:])
            if not k:
                raise IOError("MCPL: unexpected end of file")
            got += k
        self._pos += n
        return buf, n

    def device_layout(self):
        """Record layout for the device decoder (``kdsb200_mcpl_unpack``); None for
        big-endian files, which only the host decoder reads."""
        if self.hdr.endian != "<":
            return None
        from . import _lib

        f = self._dtype.fields
        lay = _lib.McplLayout()
        lay.record_bytes = self._dtype.itemsize
        lay.single_precision = int(bool(self.hdr.opt_singleprec))
        lay.off_pos = f["pos"][1]
        lay.off_dir = f["dirpack"][1]  # dirpack[2] is followed directly by ekin
        lay.off_time = f["time"][1]
        lay.off_weight = f["weight"][1] if "weight" in f else -1
        lay.off_pdg = f["pdgcode"][1] if "pdgcode" in f else -1
        lay.universal_pdg = int(self.hdr.opt_universalpdgcode)
        lay.universal_weight = float(self.hdr.opt_universalweight)
        assert f["ekin"][1] == lay.off_dir + 2 * (4 if lay.single_precision else 8)
        return lay

    def read_block(self):
        rec = self.read_raw(self.blocklength)
        if rec is None:
            return None
        return ParticleBlock(rec, self.hdr)

    @property
    def particle_blocks(self):
        while True:
            pb = self.read_block()
            if pb is None:
                return
            yield pb

    def rewind(self):
        self._fh.close()
        self._fh = _open_maybe_gz(self.filename)
        self._read_header()
        self._pos = 0

    def close(self):
        self._fh.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _pack_string(s, endian="<"):
    if isinstance(s, str):
        s = s.encode("utf8")
    return struct.pack(endian + "I", len(s)) + s


def build_header(
    nparticles,
    srcname,
    singleprec=True,
    polarisation=False,
    universal_pdg=0,
    universal_weight=0.0,
    userflags=False,
    comments=(),
):
    """Serialised MCPL-3 little-endian header."""
    dt = record_dtype(
        singleprec, polarisation, bool(universal_pdg), bool(universal_weight),
        userflags,
    )
    out = b"MCPL003L" + struct.pack("<Q", int(nparticles))
    out += struct.pack(
        "<IIIIIiII",
        len(comments),
        0,
        int(bool(userflags)),
        int(bool(polarisation)),
        int(bool(singleprec)),
        int(universal_pdg),
        dt.itemsize,
        int(bool(universal_weight)),
    )
    if universal_weight:
        out += struct.pack("<d", float(universal_weight))
    out += _pack_string(srcname)
    for c in comments:
        out += _pack_string(c)
    return out


def patch_nparticles(filename, nparticles):
    """Rewrite the particle count of an *uncompressed* MCPL file in place."""
    with open(filename, "r+b") as fh:
        fh.seek(8)
        fh.write(struct.pack("<Q", int(nparticles)))


def write_mcpl_file(
    filename,
    *,
    ekin,
    x,
    y,
    z,
    ux,
    uy,
    uz,
    time,
    weight,
    pdgcode,
    polx=None,
    poly=None,
    polz=None,
    userflags=None,
    double_prec=False,
    srcname="KDSource MCPL writer",
    compress=None,
    compresslevel=6,
):
    """Write arrays to an MCPL-3 file (host path; record layout as libmcpl).

    ``weight`` / ``pdgcode`` scalars select the universal-weight /
    universal-pdgcode header options (flags bit2 / bit1 of
    ``clib/src/writemcpl.c:4-13``).
    """
    filename = str(filename)
    if compress is None:
        compress = filename.endswith(".gz")
    rec, hdr = pack_records(ekin=ekin, x=x, y=y, z=z, ux=ux, uy=uy, uz=uz, time=time,
                            weight=weight, pdgcode=pdgcode, polx=polx, poly=poly, polz=polz,
                            userflags=userflags, double_prec=double_prec, srcname=srcname)
    with open(filename, "wb") as fh:
        if compress:
            # indexed gzip members, deflated on all cores (and read back the same way)
            for member in gz_indexed_members(hdr, compresslevel=compresslevel):
                fh.write(member)
            for member in gz_indexed_members(rec.view(np.uint8).reshape(-1),
                                             compresslevel=compresslevel):
                fh.write(member)
        else:
            fh.write(hdr)
            fh.write(rec.tobytes())
    return os.path.abspath(filename)


def pack_records(*, ekin, x, y, z, ux, uy, uz, time, weight, pdgcode, polx=None, poly=None,
                 polz=None, userflags=None, double_prec=False, srcname="KDSource MCPL writer",
                 nparticles=None):
    """(structured record array, header bytes) of an MCPL-3 file holding these particles;
    `nparticles` overrides the count written into the header (a file assembled from
    several record arrays)."""
    import numbers

    n = len(ekin)
    uni_pdg = int(pdgcode) if isinstance(pdgcode, numbers.Integral) else 0
    uni_w = float(weight) if isinstance(weight, numbers.Number) else 0.0
    has_pol = polx is not None
    dt = record_dtype(
        not double_prec, has_pol, bool(uni_pdg), bool(uni_w), userflags is not None
    )
    rec = np.zeros(n, dtype=dt)
    p0, p1, sek = pack_direction(ux, uy, uz, ekin)
    rec["pos"][:, 0] = x
    rec["pos"][:, 1] = y
    rec["pos"][:, 2] = z
    rec["dirpack"][:, 0] = p0
    rec["dirpack"][:, 1] = p1
    rec["ekin"] = sek
    rec["time"] = time
    if not uni_w:
        rec["weight"] = weight
    if not uni_pdg:
        rec["pdgcode"] = pdgcode
    if has_pol:
        rec["pol"][:, 0] = polx
        rec["pol"][:, 1] = poly
        rec["pol"][:, 2] = polz
    if userflags is not None:
        rec["userflags"] = userflags
    hdr = build_header(
        n if nparticles is None else nparticles, srcname, not double_prec, has_pol, uni_pdg,
        uni_w, userflags is not None
    )
    return rec, hdr
