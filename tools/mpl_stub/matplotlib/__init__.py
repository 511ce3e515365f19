"""Stand-in for matplotlib, only good for IMPORTING the reference's geometry / LiDAR modules.

The reference's `Utils/ObstaclesUtils.py`, `Utils/obstacles.py` and `RangeFinder/range_finder_wth_polygons_dbscan.py`
import matplotlib at module level for plotting; the arithmetic they contain needs only
`matplotlib.path.Path(vertices).contains_point(p)` (`ObstaclesUtils.py:50-57`), restated here as the even-odd crossing
test over the implicitly closed ring.  Everything else is a no-op object.  Used by bench.py's `cpu_baseline` legs that
time the reference's own functions (from baseline/_ref) and by tests/golden/make_*_golden.py; never by the product.
"""
import sys
import types


class _Any:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Any()

    def __getattr__(self, n):
        return _Any()


def _mk(name):
    m = types.ModuleType(name)
    m.__getattr__ = lambda n: _Any()
    sys.modules[name] = m
    return m


for _n in ("pyplot", "patches", "animation", "transforms", "collections", "lines"):
    setattr(sys.modules[__name__], _n, _mk("matplotlib." + _n))

_pm = types.ModuleType("matplotlib.path")


class Path:
    def __init__(self, v):
        import numpy as np
        self.v = np.asarray(v, dtype=float)

    def contains_point(self, p):
        v = self.v
        n = len(v)
        x, y = float(p[0]), float(p[1])
        inside = False
        for i in range(n):
            x1, y1 = v[i]
            x2, y2 = v[(i + 1) % n]
            if (y1 > y) != (y2 > y) and x < (x2 - x1) * (y - y1) / (y2 - y1) + x1:
                inside = not inside
        return inside


_pm.Path = Path
sys.modules["matplotlib.path"] = _pm
path = _pm
