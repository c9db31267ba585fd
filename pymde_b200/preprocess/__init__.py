"""Preprocessing: graphs, neighbours, distances, edge sampling (interface of pymde/preprocess)."""
import numpy as np
import scipy.sparse as sp
import torch

from . import data_matrix, graph, preprocess as _pre
from .graph import Graph  # noqa: F401
from .preprocess import deduplicate_edges, dissimilar_edges, sample_edges, scale  # noqa: F401


def k_nearest_neighbors(data, k, max_distance=None, verbose=False, device=None):
    """Type dispatch of pymde/preprocess/generic.py:60-108."""
    if isinstance(data, Graph):
        return graph.k_nearest_neighbors(data, k, max_distance=max_distance, verbose=verbose)
    return data_matrix.k_nearest_neighbors(data, k, max_distance=max_distance, verbose=verbose, device=device)


def distances(data, retain_fraction=1.0, verbose=False, device=None):
    """Type dispatch of pymde/preprocess/generic.py:13-57."""
    if isinstance(data, Graph):
        n = data.n_items
        g = graph.shortest_paths(data, verbose=verbose)
        if retain_fraction < 1.0:
            e, d = g.edges, g.distances
            keep = torch.randperm(e.shape[0])[: int(retain_fraction * n * (n - 1) / 2)]
            keep = torch.sort(keep).values
            return Graph.from_edges(e[keep], d[keep], n_items=n)
        return g
    return data_matrix.distances(data, retain_fraction=retain_fraction, verbose=verbose, device=device)
