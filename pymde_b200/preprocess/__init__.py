"""Preprocessing: graphs, neighbours, distances, edge sampling (interface of pymde/preprocess)."""
from . import data_matrix, generic, graph  # noqa: F401
from .generic import distances, k_nearest_neighbors  # noqa: F401
from .graph import Graph  # noqa: F401
from .preprocess import deduplicate_edges, dissimilar_edges, sample_edges, scale  # noqa: F401
