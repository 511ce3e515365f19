"""Mirror of the reference's `Utils/ObstaclesUtils.py` for the functions on the hot path.

`get_closest_point_and_normal_vector_from_obs` (reference `:60-109`) runs on the GPU (K1, csrc/halfplane.cu).
The polygon generators are host-side scenario synthesis (numpy / scipy Qhull), same parameters as the reference
(`:21-47`).
"""
import random

import numpy as np
import torch
from scipy.spatial import ConvexHull

import ldcbf_b200


def hull_ring(polygon):
    """Hull vertices in counter-clockwise order: `polygon.points[polygon.vertices]` (reference `:54`)."""
    if isinstance(polygon, ConvexHull):
        return np.ascontiguousarray(polygon.points[polygon.vertices], dtype=np.float64)
    return np.ascontiguousarray(np.asarray(polygon, dtype=np.float64))


class ObstaclesUtils:
    @staticmethod
    def set_random_seed(seed: int) -> None:
        random.seed(seed)

    @staticmethod
    def generate_circle_like_polygon(num_points: int, radius: float, center) -> ConvexHull:
        ang = np.linspace(0, 2 * np.pi, num_points)
        return ConvexHull(np.column_stack((center[0] + radius * np.cos(ang), center[1] + radius * np.sin(ang))))

    @staticmethod
    def generate_random_convex_polygon(num_points: int, x_range, y_range) -> ConvexHull:
        return ConvexHull([(random.uniform(*x_range), random.uniform(*y_range)) for _ in range(num_points)])

    @staticmethod
    def closest_points_and_normals(x, polygons):
        """(c[n,2], eta[n,2]) for all polygons at once — one K1 launch."""
        from ldcbf_b200.scenarios import pack_rings
        rings = [hull_ring(p) for p in polygons]
        if not rings:
            return np.zeros((0, 2)), np.zeros((0, 2))
        verts, nverts, nobs = pack_rings([rings])
        dev = torch.device("cuda")
        ce = ldcbf_b200.half_planes(torch.as_tensor(np.asarray(x, dtype=np.float64).reshape(1, 2), device=dev),
                                    torch.as_tensor(verts, device=dev), torch.as_tensor(nverts, device=dev),
                                    torch.as_tensor(nobs, device=dev)).cpu().numpy()[0]
        return ce[:, :2].copy(), ce[:, 2:].copy()

    @staticmethod
    def get_closest_point_and_normal_vector_from_obs(x, polygon, unitary_normal_vector: bool = False):
        """Returns (c[2,1], normal[2,1]) like the reference (`:60-109`)."""
        c, eta = ObstaclesUtils.closest_points_and_normals(x, [polygon])
        c, eta = c[0], eta[0]
        if not unitary_normal_vector:
            # un-normalised variant: (x - c), sign-flipped inside, as the reference returns it
            d = np.asarray(x, dtype=np.float64) - c
            eta = d if np.dot(d, eta) >= 0 else -d
        return c.reshape(2, 1), eta.reshape(2, 1)
