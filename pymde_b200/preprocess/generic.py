"""Type dispatch between data matrices and graphs (interface of pymde/preprocess/generic.py:13-108)."""
from . import data_matrix, graph
from .graph import Graph


def distances(data, retain_fraction=1.0, verbose=False, device=None):
    """Distances between the items of `data` (a matrix of row vectors or a Graph) as a Graph: Euclidean distances
    of all pairs / a uniform sample of them, or shortest-path lengths."""
    if isinstance(data, Graph):
        import os
        import torch
        A = data.adjacency_matrix
        # unweighted graphs: bit-parallel BFS on the device (the reference's fast path is one Cython BFS per node)
        if (torch.cuda.is_available() and A.shape[0] > 256 and bool((A.data == 1.0).all())
                and os.environ.get("PYMDE_B200_SHORTEST_PATHS", "device") != "host"):
            return graph.shortest_paths_device(data, retain_fraction=retain_fraction, device=device)
        return graph.shortest_paths(data, retain_fraction=retain_fraction, verbose=verbose)
    return data_matrix.distances(data, retain_fraction=retain_fraction, verbose=verbose, device=device)


def k_nearest_neighbors(data, k, max_distance=None, verbose=False, device=None):
    """k-nearest-neighbour graph of `data` (Euclidean for matrices, shortest-path metric for graphs)."""
    if isinstance(data, Graph):
        return graph.k_nearest_neighbors(data, k, max_distance=max_distance, verbose=verbose)
    return data_matrix.k_nearest_neighbors(data, k, max_distance=max_distance, verbose=verbose, device=device)
