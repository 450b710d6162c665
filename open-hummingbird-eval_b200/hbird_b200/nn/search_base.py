"""Plugin ABC — same contract as the reference's hbird/nn/search_base.py:3-31, so code written
against that ABC can drive the B200 backend unchanged."""
from abc import ABC, abstractmethod


class NearestNeighborSearchBase(ABC):
    """ctor(feature_memory, n_neighbors=30, distance_measure="dot_product", **kwargs) builds the
    index and adds the features; find_nearest_neighbors(q, k=None) searches it."""

    def __init__(self, feature_memory, n_neighbors=30, distance_measure="dot_product", **kwargs):
        self.feature_memory = feature_memory
        self.n_neighbors = n_neighbors
        self.distance_measure = distance_measure.lower()
        self.device = feature_memory.device
        self.index = self._initialize_index()
        self._add_features_to_index()

    @abstractmethod
    def _initialize_index(self):
        ...

    @abstractmethod
    def _add_features_to_index(self):
        ...

    @abstractmethod
    def find_nearest_neighbors(self, q, k=None):
        ...
