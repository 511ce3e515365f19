"""Mirror of `RangeFinder/range_finder_wth_polygons_dbscan.py`.

`compute_lidar_readings` (reference :26-63) runs on the GPU (K4, csrc/lidar.cu) and returns the same list of
`(x, y)` tuples / None.  The post-processing — Gaussian noise (:161-172), DBSCAN eps=0.3 min_samples=3 (:100-116),
convex hulls with a closing vertex (:65-83, :119-126) — is host code here (sklearn / scipy), the "next" row f1 of
SURVEY.md §8f.
"""
import numpy as np
import torch
from scipy.spatial import ConvexHull, QhullError
from sklearn.cluster import DBSCAN

import ldcbf_b200


def cast(lidar_position, obstacles, lidar_range, resolution=360):
    """K4 for one pose: (hit_obs[R], hit_edge[R], hit_xy[R,2]) numpy arrays.  obstacles: list of (n,2) arrays,
    edges (i, i+1 mod n) over the rows as given (reference Utils/obstacles.py:127-134)."""
    from ldcbf_b200.scenarios import pack_rings
    dev = torch.device("cuda")
    obstacles = [np.asarray(o, dtype=np.float64) for o in obstacles]
    if not obstacles:
        return (np.full(resolution, -1, np.int32), np.full(resolution, -1, np.int32), np.full((resolution, 2), np.nan))
    verts, nverts, nobs = pack_rings([obstacles])
    ho, he, xy = ldcbf_b200.lidar_cast(torch.as_tensor(np.asarray(lidar_position, dtype=np.float64).reshape(1, 2), device=dev),
                                       torch.as_tensor(verts, device=dev), torch.as_tensor(nverts, device=dev),
                                       torch.as_tensor(nobs, device=dev), float(lidar_range), int(resolution))
    return ho[0].cpu().numpy(), he[0].cpu().numpy(), xy[0].cpu().numpy()


def compute_lidar_readings(position, obstacles, lidar_range, resolution=360):
    _, _, xy = cast(position, obstacles, lidar_range, resolution)
    return [None if np.isnan(p[0]) else (float(p[0]), float(p[1])) for p in xy]


def create_convex_hull(points):
    points = np.unique(points, axis=0)
    if len(points) < 3 or np.linalg.matrix_rank(points - points[0]) < 2:
        return None
    try:
        return points[ConvexHull(points).vertices]
    except QhullError:
        return None


def retrieve_clusters(points, eps=0.3, min_samples=3):
    pts = np.array([p for p in points if p is not None])
    if pts.size == 0:
        return []
    pts = pts.reshape(-1, 2)
    labels = DBSCAN(eps=eps, min_samples=min_samples).fit(pts).labels_
    return [pts[labels == i] for i in set(labels) if i != -1]


def build_local_obstacles(clusters):
    out = []
    for cluster in clusters:
        poly = create_convex_hull(cluster)
        if poly is not None:
            out.append(np.append(poly, [poly[0]], axis=0))
    return out


def range_finder(lidar_position, obstacles, lidar_range=3.0, resolution=360, noisy=True):
    readings = compute_lidar_readings(lidar_position, obstacles, lidar_range=lidar_range, resolution=resolution)
    if noisy:
        noisy_readings = []
        for point in readings:
            if point is not None:
                noise = np.random.normal(0.0, 0.01, 2)
                noisy_readings.append((point[0] + noise[0], point[1] + noise[1]))
            else:
                noisy_readings.append(None)
        readings = noisy_readings
    clusters = retrieve_clusters(readings)
    return readings, clusters, build_local_obstacles(clusters)
