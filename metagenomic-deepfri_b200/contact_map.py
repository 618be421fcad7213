"""Drop-in for `mDeepFRI.contact_map` (`contact_map.py:6-95`): same classes and validation,
distance / contact computation on the GPU."""
import numpy as np

from .contact_map_utils import pairwise_sqeuclidean


class CAlphaCoordinates:
    def __init__(self, structure_id: str, coords: np.ndarray):
        self.structure_id = structure_id
        self.coords = coords
        if coords.shape[1] != 3:
            raise ValueError("Coordinates are not 3D.")

    def calculate_distance_map(self, distance="sqeuclidean"):
        if distance == "sqeuclidean":
            distances = pairwise_sqeuclidean(np.ascontiguousarray(self.coords.astype(np.float32)))
        else:
            raise NotImplementedError("Distance metric not implemented.")
        return DistanceMap(distances)

    def calculate_contact_map(self, threshold=6.0):
        return self.calculate_distance_map().calculate_contacts(threshold**2)


class DistanceMap:
    def __init__(self, distance_map):
        self.distance_map = distance_map
        if not np.all(distance_map >= 0):
            raise ValueError("Distance map contains negative values.")
        if not np.all(np.diag(distance_map) == 0):
            raise ValueError("Distance map diagonal is not zero.")
        if not np.allclose(distance_map, distance_map.T):
            raise ValueError("Distance map is not symmetric.")

    def calculate_contacts(self, threshold: int):
        return ContactMap((self.distance_map < threshold).astype(np.int32))


class ContactMap:
    def __init__(self, cmap):
        self.cmap = cmap
        if not np.allclose(cmap, cmap.T):
            raise ValueError("Contact map is not symmetric.")
        if not np.all(np.isin(cmap, [0, 1])):
            raise ValueError("Contact map values not in range [0, 1].")

    def sparsify(self):
        return np.argwhere(self.cmap == 1).astype(np.int32)
