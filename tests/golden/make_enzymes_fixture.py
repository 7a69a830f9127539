"""Generates tests/golden/dataset_enzymes.npz from the reference's bundled data THROUGH the reference's own loader
(load_data.read_graphfile, /root/reference/load_data.py:7-109; node order = its relabelling, isolated nodes dropped,
graphs > max_nodes dropped), so the fixture is exactly what train.py would feed (train.py:470-481: node-label
one-hot features).  Run in the build container (needs /root/reference); the GPU box only reads the .npz.

    python tests/golden/make_enzymes_fixture.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from graph_pooling_b200 import shim  # noqa: E402

shim.install('/root/reference')
import load_data  # noqa: E402  (the reference's)

graphs = load_data.read_graphfile('/root/reference/data', 'ENZYMES', max_nodes=100)
n, glabel, nlabel, eptr, edges = [], [], [], [0], []
for G in graphs:
    n.append(G.number_of_nodes())
    glabel.append(int(G.graph['label']))
    for u in range(G.number_of_nodes()):
        nlabel.append(int(np.argmax(G.node[u]['label'])))
    for u, v in G.edges():
        edges.append((min(u, v), max(u, v)))
    eptr.append(len(edges))
out = os.path.join(HERE, 'dataset_enzymes.npz')
np.savez_compressed(out, n=np.asarray(n, np.int16), glabel=np.asarray(glabel, np.int8),
                    nlabel=np.asarray(nlabel, np.int8), eptr=np.asarray(eptr, np.int32),
                    edges=np.asarray(edges, np.int16), num_node_labels=np.int32(len(graphs[0].node[0]['label'])))
print(out, len(graphs), 'graphs', sum(n), 'nodes', len(edges), 'edges', os.path.getsize(out), 'bytes')
