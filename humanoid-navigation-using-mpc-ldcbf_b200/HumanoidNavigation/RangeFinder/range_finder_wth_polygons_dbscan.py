"""Mirror of `RangeFinder/range_finder_wth_polygons_dbscan.py`.

`compute_lidar_readings` (reference :26-63) is the K4 kernel (csrc/lidar.cu); clustering and hulls
(`retrieve_clusters` :100-116, `create_convex_hull` :65-83, `build_local_obstacles` :119-126) are the f1 kernel
(csrc/lidar_clusters.cu).  The functions return the reference's Python shapes: readings as a list of `(x, y)` / None,
clusters as a list of (n,2) arrays in sklearn's label order, local obstacles as closed polygons (first vertex repeated).
The Gaussian noise of `range_finder` (:161-172) is drawn on the host from numpy's global RNG exactly like the
reference (two normals per valid reading, in ray order) and handed to the kernel.
"""
import numpy as np
import torch

import ldcbf_b200


def _dev():
    return torch.device("cuda")


def cast(lidar_position, obstacles, lidar_range, resolution=360):
    """K4 for one pose: (hit_obs[R], hit_edge[R], hit_xy[R,2]) numpy arrays.  obstacles: list of (n,2) arrays,
    edges (i, i+1 mod n) over the rows as given (reference Utils/obstacles.py:127-134)."""
    from ldcbf_b200.scenarios import pack_rings
    obstacles = [np.asarray(o, dtype=np.float64) for o in obstacles]
    if not obstacles:
        return (np.full(resolution, -1, np.int32), np.full(resolution, -1, np.int32), np.full((resolution, 2), np.nan))
    verts, nverts, nobs = pack_rings([obstacles])
    ho, he, xy = ldcbf_b200.lidar_cast(torch.as_tensor(np.asarray(lidar_position, dtype=np.float64).reshape(1, 2), device=_dev()),
                                       torch.as_tensor(verts, device=_dev()), torch.as_tensor(nverts, device=_dev()),
                                       torch.as_tensor(nobs, device=_dev()), float(lidar_range), int(resolution))
    return ho[0].cpu().numpy(), he[0].cpu().numpy(), xy[0].cpu().numpy()


def compute_lidar_readings(position, obstacles, lidar_range, resolution=360):
    _, _, xy = cast(position, obstacles, lidar_range, resolution)
    return [None if np.isnan(p[0]) else (float(p[0]), float(p[1])) for p in xy]


def _cluster(xy, eps=0.3, min_samples=3):
    """f1 for one scan given as an (R,2) array with NaN rows: (labels[R], hull rings list)."""
    R = xy.shape[0]
    out = ldcbf_b200.lidar_clusters(torch.as_tensor(np.ascontiguousarray(xy[None]), device=_dev()), eps=eps,
                                    min_samples=min_samples, max_hulls=64, max_hull_verts=max(16, min(R, 192)))
    labels = out["labels"][0].cpu().numpy()
    nv = out["nverts"][0].cpu().numpy()
    v = out["verts"][0].cpu().numpy()
    return labels, [v[o, :nv[o]].copy() for o in range(int(out["nobs"][0].item()))]


def _as_array(points):
    return np.array([[np.nan, np.nan] if p is None else [p[0], p[1]] for p in points], dtype=np.float64).reshape(-1, 2)


def retrieve_clusters(points, eps=0.3, min_samples=3):
    xy = _as_array(points)
    if xy.size == 0 or np.isnan(xy[:, 0]).all():
        return []
    labels, _ = _cluster(xy, eps, min_samples)
    return [xy[labels == i] for i in sorted(set(labels.tolist())) if i != -1]


EVERYTHING_IS_ONE_CLUSTER = 1e9      # an eps larger than any map: the hull stage of f1 alone (m; squared on the device)


def create_convex_hull(cluster):
    """Hull ring of ONE cluster (reference `:65-83`: np.unique, None for < 3 points or a flat set, else the hull vertices)
    through the hull stage of the f1 kernel: with min_samples = 1 and an eps that joins every pair the clustering stage
    returns the input as a single cluster."""
    _, rings = _cluster(np.asarray(cluster, dtype=np.float64).reshape(-1, 2), eps=EVERYTHING_IS_ONE_CLUSTER, min_samples=1)
    return rings[0] if rings else None


def build_local_obstacles(clusters):
    """Reference `:119-126`: the hull of every cluster, closed by repeating its first vertex; flat clusters are dropped."""
    hulls = (create_convex_hull(c) for c in clusters)
    return [np.append(h, [h[0]], axis=0) for h in hulls if h is not None]


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
    xy = _as_array(readings)
    if xy.size == 0 or np.isnan(xy[:, 0]).all():
        return readings, [], []
    labels, rings = _cluster(xy)
    clusters = [xy[labels == i] for i in sorted(set(labels.tolist())) if i != -1]
    local_obstacles = [np.append(r, [r[0]], axis=0) for r in rings]
    return readings, clusters, local_obstacles
