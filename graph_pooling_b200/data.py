"""Compact graph datasets and a GPU-resident batch feed (SURVEY.md 8(f) N2 + N3).

The reference turns the TU-Dortmund text files into a list of ``networkx`` graphs with Python loops
(``load_data.read_graphfile``, /root/reference/load_data.py:7-109), keeps one dense float64 ``N x N`` matrix
per graph (``graph_sampler.py:26``), pads it per item in ``__getitem__`` (:97-109), pickles the batch across the
DataLoader worker pipe and copies ``B*N*N*4`` bytes to the device every step (``train.py:197-201``).  Once the
kernels are fast that feed dominates (ENZYMES through the unchanged ``train.py``: ~9 ms per step, of which the
step itself is < 1 ms).

Here the whole dataset lives as a handful of flat arrays (``GraphSet``: node counts, labels, a CSR edge list with
the reference's node numbering), is parsed with vectorised numpy, uploaded ONCE, and every padded batch
(``x``, ``adj``, ``n_b``, ``label``) is assembled on the device from the edge list -- nothing but the graph
indices crosses PCIe per step.  ``adj`` is uint8 {0,1} for the tensor-core mode (``gp_adj_prepare`` expands it to
the bf16 operand) or float32 for the fp32 mode.

Semantics reproduced from the reference loader (checked against it in tests/test_data_cpu.py):
  * node order inside a graph = order of first appearance in ``<name>_A.txt`` (``nx.from_edgelist`` + relabel,
    load_data.py:78,96-108); nodes that never appear in an edge are dropped;
  * graphs with more than ``max_nodes`` nodes are dropped (load_data.py:79);
  * graph labels ``value - 1``, or the raw value when a label 0 exists (load_data.py:46-59);
  * node labels ``value - 1`` one-hot over ``max + 1`` classes (load_data.py:24-33,86-89); node attributes as
    float rows (load_data.py:35-44);
  * batches: symmetric {0,1} adjacency without self loops unless the file lists them, zero padding to
    ``max_nodes``, features zero on pad rows (graph_sampler.py:26-37,97-109; ``normalize=False`` as every caller
    passes it: train.py:296, cross_val.py:29).
"""
import os

import numpy as np
import torch


class GraphSet:
    """Flat-array graph dataset.  n [G] nodes per graph; label [G]; nlabel [sum n] node label ids (or None);
    attrs [sum n, A] node attributes (or None); eptr [G+1], edges [E,2]: undirected edges in graph-local ids."""

    def __init__(self, n, label, nlabel, eptr, edges, num_node_labels=0, attrs=None):
        self.n = np.asarray(n, np.int32)
        self.label = np.asarray(label, np.int64)
        self.nlabel = None if nlabel is None else np.asarray(nlabel, np.int64)
        self.attrs = None if attrs is None else np.asarray(attrs, np.float32)
        self.eptr = np.asarray(eptr, np.int64)
        self.edges = np.asarray(edges, np.int64).reshape(-1, 2)
        self.num_node_labels = int(num_node_labels)
        self.nptr = np.concatenate([[0], np.cumsum(self.n, dtype=np.int64)])
        self._dev = None

    def __len__(self):
        return len(self.n)

    @property
    def feat_dim(self):
        return self.num_node_labels if self.nlabel is not None else (0 if self.attrs is None else self.attrs.shape[1])

    # ---- device residency -----------------------------------------------------------------------------
    def to(self, device):
        """Upload the dataset once; batches are then assembled on `device`."""
        dev = torch.device(device)
        d = {'n': torch.from_numpy(self.n).to(dev), 'label': torch.from_numpy(self.label).to(dev),
             'nptr': torch.from_numpy(self.nptr).to(dev), 'eptr': torch.from_numpy(self.eptr).to(dev),
             'edges': torch.from_numpy(self.edges).to(dev),
             'egraph': torch.from_numpy(np.repeat(np.arange(len(self.n)), np.diff(self.eptr))).to(dev),
             'ngraph': torch.from_numpy(np.repeat(np.arange(len(self.n)), self.n)).to(dev),
             'nlocal': torch.from_numpy(np.concatenate([np.arange(k) for k in self.n]) if len(self.n) else
                                        np.zeros(0, np.int64)).to(dev)}
        if self.nlabel is not None:
            d['nlabel'] = torch.from_numpy(self.nlabel).to(dev)
        if self.attrs is not None:
            d['attrs'] = torch.from_numpy(self.attrs).to(dev)
        self._dev = d
        return self

    def batch(self, idx, max_nodes, adj_dtype=torch.uint8, features='node-label'):
        """Padded batch of graphs `idx` (1-D LongTensor / array) built on the dataset's device.
        Returns x [B,N,D] float32, adj [B,N,N] adj_dtype, nb [B] int32 (device), label [B] int64."""
        if self._dev is None:
            raise RuntimeError('call GraphSet.to(device) first')
        d = self._dev
        dev = d['n'].device
        idx = torch.as_tensor(idx, device=dev, dtype=torch.long)
        B, N = int(idx.numel()), int(max_nodes)
        nb = d['n'][idx]
        # gather per BATCH POSITION (a graph id may occur several times: sampling with replacement, oversampled
        # cross-validation folds): position p owns the edge range eptr[idx[p]] .. eptr[idx[p]+1] and the node range
        # nptr[idx[p]] .. nptr[idx[p]+1], expanded with repeat_interleave
        def expand(ptr):
            start = ptr[idx]
            cnt = ptr[idx + 1] - start
            pos = torch.repeat_interleave(torch.arange(B, device=dev), cnt)
            first = torch.cumsum(cnt, 0) - cnt                             # offset of each position's first item
            item = torch.arange(int(pos.numel()), device=dev) - first[pos] + start[pos]
            return pos, item

        adj = torch.zeros(B, N, N, device=dev, dtype=adj_dtype)
        eb, ei = expand(d['eptr'])
        eu, ev = d['edges'][ei, 0], d['edges'][ei, 1]
        adj[eb, eu, ev] = 1
        adj[eb, ev, eu] = 1
        nbi, nk = expand(d['nptr'])
        nli = d['nlocal'][nk]
        if features == 'node-label' and 'nlabel' in d:                     # train.py:477-481
            x = torch.zeros(B, N, self.num_node_labels, device=dev)
            x[nbi, nli, d['nlabel'][nk]] = 1.0
        elif 'attrs' in d:                                                 # feat == 'node-feat'
            x = torch.zeros(B, N, d['attrs'].shape[1], device=dev)
            x[nbi, nli] = d['attrs'][nk]
        else:
            raise ValueError('dataset has neither node labels nor node attributes')
        return x, adj, nb.to(torch.int32), d['label'][idx]


def _read_ints(path):
    with open(path) as f:
        return np.array(f.read().split(), dtype=np.int64)


def read_tu_dataset(datadir, name, max_nodes=None):
    """Vectorised reader of the TU-Dortmund text format with the semantics of the reference's
    load_data.read_graphfile (see the module docstring).  Returns a GraphSet."""
    prefix = os.path.join(datadir, name, name)
    gind = _read_ints(prefix + '_graph_indicator.txt')                      # node (1-based) -> graph (1-based)
    e = np.loadtxt(prefix + '_A.txt', delimiter=',', dtype=np.int64).reshape(-1, 2)
    glab = _read_ints(prefix + '_graph_labels.txt')
    glab = glab if (glab == 0).any() else glab - 1                          # load_data.py:46-59
    nlab = None
    num_nl = 0
    if os.path.exists(prefix + '_node_labels.txt'):
        nlab = _read_ints(prefix + '_node_labels.txt') - 1
        num_nl = int(nlab.max()) + 1
    attrs = None
    if os.path.exists(prefix + '_node_attributes.txt'):
        with open(prefix + '_node_attributes.txt') as f:
            attrs = np.array([[float(t) for t in line.replace(',', ' ').split()] for line in f], dtype=np.float32)
    G = len(glab)
    # an edge belongs to the graph of its first endpoint (load_data.py:69)
    eg = gind[e[:, 0] - 1] - 1
    # node order = first appearance in the (file-ordered) edge list of its graph
    flat = e.reshape(-1)
    fgraph = np.repeat(eg, 2)
    # a node can only appear in edges of its own graph in well-formed files; key on (graph, node) to be safe
    key = fgraph * (len(gind) + 1) + flat
    uniq, first = np.unique(key, return_index=True)
    ugraph, unode = uniq // (len(gind) + 1), uniq % (len(gind) + 1)
    order = np.lexsort((first, ugraph))
    ugraph, unode = ugraph[order], unode[order]
    counts = np.bincount(ugraph, minlength=G)
    starts = np.concatenate([[0], np.cumsum(counts)])
    local = np.arange(len(unode)) - starts[ugraph]
    # global (graph, node) -> local id lookup for the edges
    node_local = np.full(len(gind) + 1, -1, np.int64)
    node_local[unode] = local                                               # well-formed: one graph per node
    keep_g = np.ones(G, bool) if max_nodes is None else counts <= max_nodes
    keep_g &= counts > 0                                                    # nx.from_edgelist([]) has no nodes;
    # the reference keeps empty graphs only to crash later on G.node[0]; they do not occur in ENZYMES / DD
    # undirected unique edges in local ids (both directions are listed in the files; nx collapses them)
    lu, lv = node_local[e[:, 0]], node_local[e[:, 1]]
    a, b = np.minimum(lu, lv), np.maximum(lu, lv)
    ekey = (eg * (counts.max() + 1) + a) * (counts.max() + 1) + b
    _, efirst = np.unique(ekey, return_index=True)
    efirst.sort()
    eg_u, a_u, b_u = eg[efirst], a[efirst], b[efirst]
    sel = keep_g[eg_u]
    eg_u, a_u, b_u = eg_u[sel], a_u[sel], b_u[sel]
    eorder = np.argsort(eg_u, kind='stable')
    eg_u, a_u, b_u = eg_u[eorder], a_u[eorder], b_u[eorder]
    new_id = np.cumsum(keep_g) - 1
    ecount = np.bincount(new_id[eg_u], minlength=int(keep_g.sum()))
    eptr = np.concatenate([[0], np.cumsum(ecount)])
    nsel = keep_g[ugraph]
    nodes_kept = unode[nsel]
    return GraphSet(counts[keep_g], glab[keep_g], None if nlab is None else nlab[nodes_kept - 1], eptr,
                    np.stack([a_u, b_u], 1), num_nl, None if attrs is None else attrs[nodes_kept - 1])
