"""Import stub: the reference imports matplotlib (usetex) at module import on the
geometry path (plotTools/pM.py:1-3, plotResults.py:1-3); nothing on the hot path uses it."""
rcParams = {}


def rc(*a, **k):
    pass
