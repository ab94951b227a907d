"""Small helpers shared by the package (reference python/kdsource/utils.py:12-40,
clib/src/utils.c:106-125)."""

_PT2PDG = {"n": 2112, "p": 22, "e": 11}
_PDG2PT = {v: k for k, v in _PT2PDG.items()}


def pt2pdg(pt):
    """Particle-type character -> PDG code (0 if unknown)."""
    return _PT2PDG.get(pt, 0)


def pdg2pt(pdgcode):
    """PDG code -> particle-type character ("0" if unknown)."""
    return _PDG2PT.get(int(pdgcode), "0")
