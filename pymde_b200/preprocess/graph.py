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

    def draw(self, embedding_dim=2, standardized=False, device=None, verbose=False):
        """Embed the graph for drawing (pymde/preprocess/graph.py:200-256): shortest-path distances (at most 1e7
        of them), WeightedQuadratic loss + Centered, or Cubic penalty + Standardized.  Returns the embedding;
        plotting is out of scope here."""
        from .. import constraints, problem
        from ..functions import losses, penalties
        if bool((self.distances < 0).any()):
            raise ValueError("Graphs with negative edge weights cannot be drawn.")
        if self.n_edges < self.n_all_edges:
            dg = shortest_paths(self, retain_fraction=min(1.0, 1e7 / self.n_all_edges), verbose=verbose)
        else:
            dg = self
        if not standardized:
            constraint, f = constraints.Centered(), losses.WeightedQuadratic(dg.distances)
        else:
            constraint, f = constraints.Standardized(), penalties.Cubic(1 / dg.distances)
        mde = problem.MDE(n_items=self.n_items, embedding_dim=embedding_dim, edges=dg.edges, distortion_function=f,
                          constraint=constraint, device=device)
        return mde.embed(verbose=verbose)

    def __getitem__(self, key):
        return self._A[key]

    def __setitem__(self, key, value):
        raise AttributeError("Graph objects are immutable.")


class EdgeListGraph(object):
    """Result of the device shortest-path search: an edge list that never left the GPU.  Quacks like the part of
    `Graph` the recipes use (`edges` sorted by (i, j), `distances`, `n_items`); `to_graph()` builds the scipy-backed
    object when somebody needs the adjacency matrix."""

    def __init__(self, edges, distances, n_items):
        self.edges, self.distances, self._n = edges, distances, int(n_items)
        self.weights = distances

    @property
    def n_items(self):
        return self._n

    @property
    def n_edges(self):
        return int(self.edges.shape[0])

    @property
    def n_all_edges(self):
        return self._n * (self._n - 1) // 2

    def to_graph(self):
        return Graph.from_edges(self.edges.cpu().numpy(), self.distances.cpu().numpy(), n_items=self._n)


def shortest_paths_device(graph, max_length=None, retain_fraction=1.0, device=None, seed=None):
    """Hop-count shortest paths of an UNWEIGHTED graph on the GPU (`mde_graph_hops`: bit-parallel multi-source BFS,
    256 sources per pass over the adjacency).  Same contract as `shortest_paths` -- pairs (i < j) within `max_length`
    hops, each kept with probability `retain_fraction` -- but the sample is drawn by a counter-based hash seeded from
    the module RNG, and the result stays on the device as an `EdgeListGraph`."""
    import ctypes as C
    from .. import _lib, util
    A = graph.adjacency_matrix if isinstance(graph, Graph) else Graph(graph).adjacency_matrix
    n = A.shape[0]
    dev = util.cuda_device(device)
    lib = _lib.load()
    indptr = torch.tensor(A.indptr.astype(np.int32), device=dev)
    indices = torch.tensor(A.indices.astype(np.int32), device=dev)
    ws = torch.empty(int(lib.mde_graph_hops_ws_bytes(n)), dtype=torch.uint8, device=dev)
    if seed is None:
        seed = int(util.np_rng().integers(0, 2 ** 62))
    expected = min(1.0, float(retain_fraction)) * n * (n - 1) / 2
    cap = int(expected * 1.02 + 4 * (expected ** 0.5) + 1024)
    limit = 0 if (max_length is None or not np.isfinite(max_length)) else int(max_length)
    while True:
        src = torch.empty(cap, dtype=torch.int32, device=dev)
        dst = torch.empty(cap, dtype=torch.int32, device=dev)
        ln = torch.empty(cap, dtype=torch.float32, device=dev)
        count = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.mde_graph_hops(indptr.data_ptr(), indices.data_ptr(), n, 0, n, limit, float(retain_fraction),
                                          C.c_uint64(seed), src.data_ptr(), dst.data_ptr(), ln.data_ptr(), cap,
                                          count.data_ptr(), ws.data_ptr(), ws.numel(), util.stream_ptr(dev)))
        got = int(count.item())
        if got <= cap:
            break
        cap = int(got * 1.01) + 1024  # (only when the estimate was exceeded: same seed => same sample)
    src, dst, ln = src[:got].long(), dst[:got].long(), ln[:got]
    order = torch.argsort(src * n + dst)
    return EdgeListGraph(torch.stack([src[order], dst[order]], 1), ln[order], n)


def shortest_paths(graph, max_length=None, retain_fraction=1.0, n_workers=None, verbose=False):
    """Shortest-path distances as a Graph (interface of pymde/preprocess/graph.py:345-474): unreachable pairs and
    pairs beyond `max_length` are dropped; with `retain_fraction` < 1 every remaining pair is kept with that
    probability (Bernoulli draws from the module RNG seeded by `pymde_b200.seed`, like the reference's per-row
    sampling), which bounds memory on large graphs.  Unweighted graphs use BFS hop counts, weighted ones Dijkstra
    (scipy.sparse.csgraph), in row chunks."""
    from .. import util
    del n_workers
    if sp.issparse(graph):
        graph = Graph(graph)
    elif not isinstance(graph, Graph):
        raise ValueError("`graph` must be a pymde.Graph instance or scipy.sparse adjacency matrix.")
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
        if retain_fraction < 1.0:
            keep &= util.np_rng().uniform(size=keep.size) <= retain_fraction
        rows.append(idx[r][keep]); cols.append(c[keep]); vals.append(D[r, c][keep])
    rows, cols, vals = np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
    return Graph.from_edges(np.stack([rows, cols], 1), vals.astype(np.float32), n_items=n)


def scale(graph, natural_length):
    """New graph whose distances have RMS `natural_length` (pymde/preprocess/graph.py:259-279)."""
    d = graph.distances
    alpha = float(natural_length) / float(d.pow(2).mean().sqrt())
    return Graph.from_edges(graph.edges, alpha * d, n_items=graph.n_items)


def breadth_first_order(csgraph_matrix, i_start, directed=True, return_predecessors=True):
    """Hop counts from `i_start` (inf where unreachable) and BFS predecessors (-9999 where none), the pair the
    reference's Cython helper returns (pymde/preprocess/graph.py:286-308)."""
    del return_predecessors
    order, pred = csgraph.breadth_first_order(csgraph_matrix, i_start, directed=directed, return_predecessors=True)
    n = csgraph_matrix.shape[0]
    lengths = np.full(n, np.inf, dtype=np.float32)
    lengths[i_start] = 0.0
    for v in order[1:]:  # BFS order: a node's predecessor is always settled before the node
        lengths[v] = lengths[pred[v]] + 1.0
    pred = pred.astype(np.int32)
    pred[pred < 0] = -9999
    return lengths, pred


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
