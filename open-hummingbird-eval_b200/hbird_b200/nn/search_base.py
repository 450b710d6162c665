"""The plugin contract every nn_method backend honours.

It is the same contract as the reference's abstract base (hbird/nn/search_base.py:3-31): a backend
is constructed from the (N, d) feature memory plus `n_neighbors` / `distance_measure`, builds its
index, ingests the features, and then answers `find_nearest_neighbors(q, k=None)` with a pair of
host arrays.  Code written against the reference ABC can therefore drive the B200 backend as is.
"""
import abc


class NearestNeighborSearchBase(metaclass=abc.ABCMeta):
    #: the three hooks a backend provides, in the order the constructor drives them
    HOOKS = ("_initialize_index", "_add_features_to_index", "find_nearest_neighbors")

    def __init__(self, feature_memory, n_neighbors=30, distance_measure="dot_product", **kwargs):
        self._remember(feature_memory, int(n_neighbors), str(distance_measure))
        self.index = self._initialize_index()   # 1. empty index
        self._add_features_to_index()           # 2. ingest the bank

    def _remember(self, feature_memory, n_neighbors, distance_measure):
        self.feature_memory, self.device = feature_memory, feature_memory.device
        self.n_neighbors, self.distance_measure = n_neighbors, distance_measure.lower()

    @abc.abstractmethod
    def _initialize_index(self):
        """Return the (still empty) index object."""

    @abc.abstractmethod
    def _add_features_to_index(self):
        """Move `self.feature_memory` into `self.index`."""

    @abc.abstractmethod
    def find_nearest_neighbors(self, q, k=None):
        """(indices (Q, k), distances (Q, k)) of the k best bank rows per query row of `q`."""
