"""``resample_to_mcpl``: draw new particles from an XML source file on the GPU.

Same call as the reference (``python/kdsource/resample.py:2-49``); the C loop
behind it (``kdsource_resample_to_mcpl``, clib/src/resamplemcpl.c:17-47) is
replaced by the batched device sampler.  The XML file is read the way
``KDS_open`` reads it (clib/src/kdsource.c:79-280): J, optional kernel, PList,
Geom with its metrics, scaling, BW (scalar, or a raw float32 file with one
value per valid particle).
"""

import os
import pathlib
import shutil
import zlib
from xml.etree.ElementTree import parse

import numpy as np

from . import _lib, mcpl
from .geom import Geometry
from .plist import PList
from .sampler import DEFAULT_SLOTS, DeviceSampler, make_geom_struct
from .utils import pt2pdg
from .writemcpl import _check_new_mcpl_filename, _normalised_mcpl_filename


def open_source(kds_sourcefile, precision="float32", contract=2):
    """Parse an XML source file and upload its particle list: returns a
    :class:`DeviceSampler` plus a dict of the parsed fields."""
    path = pathlib.Path(kds_sourcefile)
    root = parse(path).getroot()
    if root.tag != "KDSource":
        raise RuntimeError("Invalid format in source XML file {}".format(path))
    kel = root.find("kernel")
    kernel = (kel.text or "g").strip()[0] if kel is not None else "g"
    pltree = root.find("PList")
    # relative MCPL names are resolved against the XML file's directory first
    name = pltree[1].text
    if not os.path.isabs(name) and not os.path.isfile(name):
        cand = path.parent / name
        if cand.is_file():
            pltree[1].text = str(cand)
    cwd = os.getcwd()
    plist = PList.load(pltree, set_params=False)
    geom = Geometry.load(root.find("Geom"))
    scaling = np.array(root.find("scaling").text.split(), dtype="float64")
    bwel = root.find("BW")
    variable = bool(int(bwel.attrib["variable"]))
    parts, ws = plist.get_device(-1)  # decoded on the device, stays there
    if len(parts) == 0:
        raise RuntimeError("Empty particle list")
    if variable:
        bw = np.fromfile(bwel.text, dtype="float32")
        if len(bw) != len(parts):
            print("Warning: Particle list and bandwidths file have different size.")
        if len(bw) < len(parts):
            bw = np.resize(bw, len(parts))  # Geom_next rewinds the file, geom.c:160-177
    else:
        bw = float(bwel.text)
    gs = make_geom_struct(geom, scaling, kernel)
    sampler = DeviceSampler(parts, ws, bw, gs, pdgcode=pt2pdg(plist.pt), precision=precision,
                            contract=contract)
    info = dict(J=float(root.find("J").text), kernel=kernel, plist=plist, geom=geom,
                scaling=scaling, bw=bw, N=len(parts), cwd=cwd)
    return sampler, info


def _gzip_member(data, level):
    """One complete gzip member (concatenated members form a valid .gz file, which
    gzip/zlib readers -- and libmcpl's gzread -- decode as one stream)."""
    co = zlib.compressobj(level, zlib.DEFLATED, 31)
    return co.compress(data) + co.flush()


def resample_to_mcpl(kds_sourcefile, destination_mcpl, nout, rng=None, force=False,
                     batch=None, compresslevel=None, precision="float32", threads=None,
                     contract=2, one_file=True):
    """Resample `nout` particles from an XML source file into a new MCPL file
    (always ``*.mcpl.gz``; refuses to overwrite unless `force`).

    rng: an int seed (or None: seeded from os.urandom) keys the counter-based
    Philox generator, so output particle i depends only on (seed, i); a callable
    returning uniforms in [0,1) is honoured too (it is consumed on the host,
    32 numbers per particle, and fed to the same kernel).

    compresslevel=None (default): the MCPL-3 records are gzip-compressed on the
    GPU (``kdsb200_gz_writer_*``: order-0 Huffman deflate members), copied to
    pinned buffers and written by a native worker thread while the next batch is
    generated.  An int selects host zlib at that level instead (`threads` worker
    threads, default all cores; 0 = stored).  With torch.distributed initialised
    each rank produces a contiguous share of the particles and rank 0 assembles
    the file; the particles do not depend on the number of ranks.  ``one_file=False``
    skips the assembly: rank r > 0 leaves its share as a complete MCPL file of its own,
    ``<name>.rank<r>.mcpl.gz`` (all ranks write concurrently, nothing is copied).

    contract: how Philox words become Gaussian deviates -- 2 (default) Box-Muller pairs,
    1 the reference's polar method (``kdsb200_resample`` in include/kdsb200.h); the
    distribution is the same, the particles of a given seed are not.
    """
    from . import dist as _dist

    ksrc = pathlib.Path(kds_sourcefile)
    nout = int(nout)
    if not ksrc.is_file():
        raise RuntimeError(f"Missing file: {kds_sourcefile}")
    rank, world = _dist.rank_world()
    # argument checks first: nothing on disk is touched before they have passed
    if nout <= 0 or nout > 1e16:
        raise RuntimeError("Invalid (or unreasonable) number of"
                           f" particles requested: {nout}")
    if world > 1 and not isinstance(rng, int):
        raise RuntimeError("multi-rank resampling needs an explicit integer seed")
    if rng is None:
        rng = int.from_bytes(os.urandom(8), byteorder="big")
        print("Warning: No explicit seed provided"
              f" (choosing arbitrarily seed={rng})")
    # Only rank 0 looks at (and, with force, removes) an existing output file; its
    # verdict is shared so that a refusal raises on every rank instead of leaving the
    # others waiting at a barrier.
    refusal = None
    if rank == 0:
        try:
            dst = _check_new_mcpl_filename(destination_mcpl, force=force)
        except RuntimeError as e:
            refusal = e
    else:
        try:
            dst = _normalised_mcpl_filename(destination_mcpl).absolute()
        except RuntimeError as e:
            refusal = e
    if not _dist.all_agree(refusal is None):
        raise refusal if refusal is not None else RuntimeError(
            "resample_to_mcpl: rank 0 refused the output file " + str(destination_mcpl))
    import time

    timing = os.environ.get("KDSB200_TIMING")
    t_start = time.perf_counter()
    sampler, _ = open_source(ksrc.absolute(), precision=precision, contract=contract)
    if timing:
        print("[timing] open_source {:.3f} s".format(time.perf_counter() - t_start))
    lo, hi = _dist.shard_range(nout, rank, world)
    part_name = str(dst) + ".part{}".format(rank) if world > 1 else None
    header = mcpl.build_header(nout, "KDSource resample", singleprec=True) if rank == 0 else None
    if world > 1 and not one_file:  # one complete file per rank
        if rank > 0:
            part_name = str(dst)[:-len(".mcpl.gz")] + ".rank{}.mcpl.gz".format(rank)
        else:
            part_name = str(dst)
        header = mcpl.build_header(hi - lo, "KDSource resample", singleprec=True)
    if batch is None:
        # small batches keep the pinned staging buffers (and their allocation time) small;
        # long runs amortise larger ones better (measured: 1e8 particles 0.62 s with 2^20,
        # 1e9 particles 4.8 s with 2^22)
        batch = 1 << 22 if hi - lo > (1 << 27) else 1 << 20
    if not isinstance(rng, int):
        batch = min(batch, 1 << 18)  # host-generated uniforms: 256 B per particle
    batch = int(max(1, min(batch, hi - lo)))
    print("Resampling...")

    def batches():
        """(device record tensor, count) per batch; the tensor is reused."""
        t = _lib.torch()
        rec = _lib.empty(batch * 36, t.uint8)
        done = lo
        while done < hi:
            n = min(batch, hi - done)
            if isinstance(rng, int):
                sampler.sample_device(n, seed=rng, first=done, out=rec)
            else:
                ext = np.fromiter((rng() for _ in range(n * DEFAULT_SLOTS)), dtype=np.float64,
                                  count=n * DEFAULT_SLOTS).reshape(n, DEFAULT_SLOTS)
                sampler.sample_device(n, first=done, ext_uniforms=ext, out=rec)
            yield rec, n
            done += n

    if compresslevel is None:
        _write_gpu_gzip(part_name or str(dst), header, batches(), batch)
    else:
        _write_host_gzip(part_name or str(dst), header, batches(), compresslevel, threads)
    if timing:
        print("[timing] total {:.3f} s".format(time.perf_counter() - t_start))
    if world > 1 and not one_file:
        _dist.barrier()
    elif world > 1:
        _dist.barrier()
        if rank == 0:
            with open(dst, "wb") as out:
                for r in range(world):
                    pn = str(dst) + ".part{}".format(r)
                    with open(pn, "rb") as src:
                        shutil.copyfileobj(src, out, 1 << 24)
                    os.unlink(pn)
        _dist.barrier()
    if rank == 0:
        print("Successfully sampled {} particles.".format(nout))


def _write_gpu_gzip(path, header, batches, batch):
    """Device-compressed output through the native streaming writer."""
    import ctypes

    import time

    L = _lib.load()
    timing = os.environ.get("KDSB200_TIMING")
    t0 = time.perf_counter()
    # MCPL-3 records: 36 bytes, the last 12 (time, weight, pdgcode) usually repeat
    w = L.kdsb200_gz_writer_open(os.fsencode(path), batch * 36, 36, 12, _lib.stream())
    if not w:
        _lib.check(1)
    w = ctypes.c_void_p(w)
    if timing:
        print("[timing] gz_writer_open {:.3f} s".format(time.perf_counter() - t0))
    try:
        if header is not None:
            _lib.check(L.kdsb200_gz_writer_put_host(w, header, len(header)))
        for rec, n in batches:
            _lib.check(L.kdsb200_gz_writer_put_device(w, _lib.ptr(rec), n * 36))
    except BaseException:
        L.kdsb200_gz_writer_close(w, None)
        raise
    if timing:
        print("[timing] batches submitted {:.3f} s".format(time.perf_counter() - t0))
    nbytes = ctypes.c_uint64(0)
    _lib.check(L.kdsb200_gz_writer_close(w, ctypes.byref(nbytes)))
    if timing:
        print("[timing] writer closed {:.3f} s".format(time.perf_counter() - t0))
    return int(nbytes.value)


def _write_host_gzip(path, header, batches, compresslevel, threads):
    """Host zlib output: independent gzip members compressed on a thread pool while
    the next batch is generated."""
    from concurrent.futures import ThreadPoolExecutor

    nthreads = threads or os.cpu_count() or 1
    with open(path, "wb") as fh, ThreadPoolExecutor(nthreads) as pool:
        if header is not None:
            fh.write(_gzip_member(header, compresslevel))
        pending = []  # futures of the previous batch, written while the next one is generated

        def drain():
            for fut in pending:
                fh.write(fut.result())
            pending.clear()

        for rec, n in batches:
            raw = rec[:n * 36].cpu().numpy()
            drain()
            step = max(36 * 4096, (len(raw) // nthreads + 35) // 36 * 36)
            for s0 in range(0, len(raw), step):
                pending.append(pool.submit(_gzip_member, raw[s0:s0 + step].tobytes(),
                                           compresslevel))
        drain()
