"""Weighted graphs for MDE problem construction (interface of pymde/preprocess/graph.py:75-256).

A `Graph` wraps a symmetric scipy CSR adjacency matrix.  `edges` lists each undirected edge once,
(i, j) with i < j, sorted by (i, j); `distances` / `weights` are the matching values.  Problem
construction is one-shot host work (outside the hot path); the resulting
edge tensors are what the CUDA path consumes."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.csgraph as csgraph
import torch


def _as_numpy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


class Graph(object):
    def __init__(self, adjacency_matrix):
        A = adjacency_matrix
        if isinstance(A, torch.Tensor):
            A = A.detach().cpu().numpy()
        if isinstance(A, np.ndarray):
            A = sp.csr_matrix(A)
        elif sp.issparse(A) and not isinstance(A, sp.csr_matrix):
            A = A.tocsr()
        elif not sp.issparse(A):
            raise ValueError("adjacency_matrix must be a dense array, tensor or scipy sparse matrix")
        A = A.copy()
        A.data[A.data == np.inf] = 0  # unreachable pairs carry no edge
        A.eliminate_zeros()
        diag = A.diagonal() > 0
        if diag.any():
            raise ValueError("Adjacency matrices must not contain self edges; the following nodes were "
                             "found to have self edges: ", np.argwhere(diag).flatten())
        self._A = A
        self._edges = None
        self._values = None

    @staticmethod
    def from_edges(edges, weights=None, n_items=None):
        """Graph from an edge list; repeated (or reciprocal) edges have their weights summed."""
        e = _as_numpy(edges).astype(np.int64).copy()
        w = np.ones(e.shape[0], dtype=np.float32) if weights is None else _as_numpy(weights).astype(np.float32)
        lo, hi = np.minimum(e[:, 0], e[:, 1]), np.maximum(e[:, 0], e[:, 1])
        n = int(hi.max()) + 1 if n_items is None else int(n_items)
        upper = sp.coo_matrix((w, (lo, hi)), shape=(n, n)).tocsr()  # duplicates are summed here
        return Graph((upper + upper.T).tocsr())

    # --- adjacency -----------------------------------------------------------------------
    @property
    def adjacency_matrix(self):
        return self._A

    A = adjacency_matrix

    @property
    def n_items(self):
        return self._A.shape[0]

    @property
    def n_all_edges(self):
        return self.n_items * (self.n_items - 1) // 2

    def _materialise(self):
        if self._edges is None:
            U = sp.triu(self._A, k=1, format="csr").tocoo()  # row-major => sorted by (i, j)
            order = np.lexsort((U.col, U.row))
            self._edges = torch.tensor(np.stack([U.row[order], U.col[order]], 1).astype(np.int64))
            self._values = torch.tensor(U.data[order].astype(np.float32))

    @property
    def edges(self):
        self._materialise()
        return self._edges

    @property
    def distances(self):
        self._materialise()
        return self._values

    weights = distances

    @property
    def n_edges(self):
        return int(self.edges.shape[0])

    def neighbors(self, node):
        return self._A.indices[self._A.indptr[node]:self._A.indptr[node + 1]]

    def neighbor_distances(self, node):
        return self._A.data[self._A.indptr[node]:self._A.indptr[node + 1]]

    def __getitem__(self, key):
        return self._A[key]

    def __setitem__(self, key, value):
        raise AttributeError("Graph objects are immutable.")


def shortest_paths(graph, max_length=None, n_workers=None, verbose=False):
    """All-pairs shortest-path distances as a Graph (unreachable / beyond max_length pairs dropped).
    Unweighted graphs use BFS hop counts, weighted ones Dijkstra (scipy.sparse.csgraph)."""
    A = graph.adjacency_matrix
    unweighted = bool((A.data == 1.0).all())
    limit = np.inf if max_length is None else float(max_length)
    n = graph.n_items
    rows, cols, vals = [], [], []
    chunk = max(1, min(n, int(2e7 // max(n, 1))))
    for s0 in range(0, n, chunk):
        idx = np.arange(s0, min(n, s0 + chunk))
        D = csgraph.dijkstra(A, directed=False, indices=idx, unweighted=unweighted, limit=limit)
        r, c = np.nonzero(np.isfinite(D) & (D > 0))
        keep = c > idx[r]
        rows.append(idx[r][keep]); cols.append(c[keep]); vals.append(D[r, c][keep])
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    return Graph.from_edges(np.stack([rows, cols], 1), vals.astype(np.float32), n_items=n)


def k_nearest_neighbors(graph, k, graph_distances=False, max_distance=None, verbose=False):
    """k nearest neighbours of every node under the shortest-path metric."""
    A = graph.adjacency_matrix
    n = graph.n_items
    unweighted = bool((A.data == 1.0).all())
    limit = np.inf if max_distance is None else float(max_distance)
    src, dst, val = [], [], []
    chunk = max(1, min(n, int(2e7 // max(n, 1))))
    for s0 in range(0, n, chunk):
        idx = np.arange(s0, min(n, s0 + chunk))
        D = csgraph.dijkstra(A, directed=False, indices=idx, unweighted=unweighted, limit=limit)
        D[np.arange(len(idx)), idx] = np.inf
        kk = min(k, n - 1)
        nb = np.argpartition(D, kk - 1, axis=1)[:, :kk]
        dd = np.take_along_axis(D, nb, 1)
        ok = np.isfinite(dd)
        src.append(np.repeat(idx, kk)[ok.ravel()]); dst.append(nb.ravel()[ok.ravel()]); val.append(dd.ravel()[ok.ravel()])
    e = np.stack([np.concatenate(src), np.concatenate(dst)], 1)
    if graph_distances:
        lo, hi = np.minimum(e[:, 0], e[:, 1]), np.maximum(e[:, 0], e[:, 1])
        key, first = np.unique(lo * n + hi, return_index=True)
        return Graph.from_edges(np.stack([key // n, key % n], 1), np.concatenate(val)[first].astype(np.float32), n)
    return Graph.from_edges(e, None, n_items=n)
