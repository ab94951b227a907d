"""``write_mcpl``: numpy arrays -> MCPL file (reference
``python/kdsource/writemcpl.py`` + ``clib/src/writemcpl.c:15-69``)."""

import numbers
import pathlib

from . import mcpl


def _normalised_mcpl_filename(filename):
    """``*.mcpl`` -> ``*.mcpl.gz``; any other ending is an error.  No side effects."""
    dst = pathlib.Path(filename)
    if dst.name.endswith(".mcpl"):
        dst = dst.parent.joinpath(dst.name + ".gz")
    elif not dst.name.endswith(".mcpl.gz"):
        raise RuntimeError("Output name does not end with .mcpl or"
                           f" .mcpl.gz: {dst.name}")
    return dst


def _check_new_mcpl_filename(filename, force=False):
    """Output-name policy of the reference (writemcpl.py:52-74): the file is
    always ``*.mcpl.gz``; existing ``.mcpl`` / ``.mcpl.gz`` files are an error
    unless `force`; the directory must exist."""
    dst = _normalised_mcpl_filename(filename)
    for cand in (dst.parent.joinpath(dst.name[:-3]), dst):
        if cand.is_file():
            if not force:
                raise RuntimeError(f"Already exist: {cand}")
            cand.unlink()
    if not dst.parent.is_dir():
        raise RuntimeError(f"Directory not found: {dst.parent}")
    return dst.absolute()


def write_mcpl(filename, *, ekin, x, y, z, ux, uy, uz, time, weight, pdgcode,
               polx=None, poly=None, polz=None, userflags=None,
               double_prec=False, force=False):
    """Create an MCPL file from per-particle arrays.  `weight` / `pdgcode` may be
    scalars (stored once in the header); polarisation and userflags are optional;
    `double_prec` stores doubles.  The file name always ends in ``.mcpl.gz``."""
    lens = {len(a) for a in (ekin, x, y, z, ux, uy, uz, time)}
    npol = sum(p is None for p in (polx, poly, polz))
    if npol not in (0, 3):
        raise RuntimeError("Please provide all or none of polx, poly, and polz")
    for a in (polx, poly, polz, userflags):
        if a is not None:
            lens.add(len(a))
    if not isinstance(weight, numbers.Number):
        lens.add(len(weight))
    if not isinstance(pdgcode, numbers.Integral):
        lens.add(len(pdgcode))
    if len(lens) != 1:
        raise RuntimeError("Not all provided data arrays have the same length")
    n = int(lens.pop())
    if n < 0 or n > 1e16:
        raise RuntimeError(f"Invalid (or unreasonable) number of particles requested: {n}")
    dst = _check_new_mcpl_filename(filename, force=force)
    mcpl.write_mcpl_file(
        dst, ekin=ekin, x=x, y=y, z=z, ux=ux, uy=uy, uz=uz, time=time, weight=weight,
        pdgcode=pdgcode, polx=polx, poly=poly, polz=polz, userflags=userflags,
        double_prec=double_prec, srcname="KDSource MCPL writer", compress=True)
