"""Oracle: LiDAR post-processing and the unknown-environment half-planes.  TEST INFRASTRUCTURE ONLY.

Restates `/root/reference/HumanoidNavigation/RangeFinder/range_finder_wth_polygons_dbscan.py`:
* `:100-116` `retrieve_clusters` — the reference calls `sklearn.cluster.DBSCAN(eps=0.3, min_samples=3)`; `clusters`
  here does the same call (sklearn is the reference's own dependency, `requirements.txt:8`), and `dbscan_labels` is an
  independent numpy restatement of what that call computes, so the two can be checked against each other;
* `:65-83` `create_convex_hull`, `:119-126` `build_local_obstacles` — same numpy / scipy calls as the reference;
and `MPC/HumanoidMPCVariants/HumanoidMPCUnknownEnvironment.py:30-68` (`_get_list_c_and_eta` of the variant).
"""
import numpy as np
from scipy.spatial import ConvexHull, QhullError

from . import halfplane, lidar


def dbscan_labels(points, eps=0.3, min_samples=3):
    """Labels sklearn's DBSCAN assigns (restated): core = >= min_samples points within eps (self included); clusters =
    connected components of core points numbered by their smallest core index; a border point takes the
    lowest-numbered cluster among its core neighbours; -1 otherwise."""
    P = len(points)
    if P == 0:
        return np.zeros(0, dtype=int)
    d2 = ((points[:, None, :] - points[None, :, :]) ** 2).sum(-1)
    adj = d2 <= eps * eps
    core = adj.sum(1) >= min_samples
    lab = np.where(core, np.arange(P), P + 1)
    while True:
        new = lab.copy()
        for i in np.nonzero(core)[0]:
            nb = np.nonzero(adj[i] & core)[0]
            new[i] = min(lab[i], lab[nb].min())
        if np.array_equal(new, lab):
            break
        lab = new
    roots = sorted(set(lab[core]))
    cid = {r: k for k, r in enumerate(roots)}
    out = np.full(P, -1, dtype=int)
    for i in range(P):
        if core[i]:
            out[i] = cid[lab[i]]
        else:
            nb = np.nonzero(adj[i] & core)[0]
            if len(nb):
                out[i] = cid[lab[nb].min()]
    return out


def clusters(readings, eps=0.3, min_samples=3):
    """`retrieve_clusters` (`:100-116`).  readings: (R,2) array with NaN rows for missing readings."""
    from sklearn.cluster import DBSCAN
    pts = readings[~np.isnan(readings[:, 0])]
    if pts.size == 0:
        return [], np.zeros(0, dtype=int)
    labels = DBSCAN(eps=eps, min_samples=min_samples).fit(pts).labels_
    return [pts[labels == i] for i in sorted(set(labels)) if i != -1], labels


def hull_of(cluster):
    """`create_convex_hull` (`:65-83`): unique points, None for < 3 points or collinear sets, else hull vertices (CCW)."""
    pts = np.unique(cluster, axis=0)
    if len(pts) < 3 or np.linalg.matrix_rank(pts - pts[0]) < 2:
        return None
    try:
        return pts[ConvexHull(pts).vertices]
    except QhullError:            # `:81-83`: points flat to Qhull's own roundoff bound (a few 1e-15 here)
        return None


def local_obstacles(readings, eps=0.3, min_samples=3):
    """Hull vertex rings (CCW, no closing vertex) of the clusters the reference keeps, in cluster order."""
    cl, _ = clusters(readings, eps, min_samples)
    return [h for h in (hull_of(c) for c in cl) if h is not None]


def unknown_env_half_planes(position, obstacles_points, lidar_range, resolution=360, noise=None):
    """`HumanoidMPCUnknownEnvironment._get_list_c_and_eta` (`:30-68`) with injected noise: scan -> clusters -> hulls ->
    (c, eta) per inferred hull.  Returns (c[n,2], eta[n,2], rings, readings)."""
    _, _, xy = lidar.cast(position, obstacles_points, lidar_range, resolution)
    if noise is not None:
        xy = xy + np.where(np.isnan(xy), 0.0, noise)
    rings = local_obstacles(xy)
    c, eta = halfplane.half_planes(np.asarray(position, dtype=np.float64), rings)
    return c, eta, rings, xy
