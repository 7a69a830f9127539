"""CPU: plumbing of graph_pooling_b200/shim.py -- the reference's train.py (byte-unchanged) imports and runs one
epoch under the shim's compatibility layer (networkx API, matplotlib / tensorboardX / community stand-ins, argv).
The CUDA encoders cannot run here, so THIS TEST swaps the oracle in as `encoders` and makes `.cuda()` a no-op;
the product launcher always installs graph_pooling_b200.encoders (asserted below).  Needs the read-only
reference checkout, which exists in the build container only: skipped elsewhere."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'

SCRIPT = r'''
import sys
sys.path.insert(0, %(root)r)
import torch
from graph_pooling_b200 import shim
stubbed = shim.install(%(ref)r, seed=0)
import encoders
assert encoders.__name__ == 'graph_pooling_b200.encoders', encoders.__name__   # what the product launcher installs
assert encoders.SoftPoolingGcnEncoder.__init__.__code__.co_varnames[:7] == \
    ('self', 'max_num_nodes', 'input_dim', 'hidden_dim', 'embedding_dim', 'label_dim', 'num_layers')
# ---- test-only substitution: CPU oracle + no-op .cuda() ----
from oracle import diffpool_oracle
sys.modules['encoders'] = diffpool_oracle
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
import os
os.makedirs('results', exist_ok=True)
sys.argv = ['train.py', '--bmname=ENZYMES', '--datadir=%(ref)s/data', '--method=%(method)s', '--max-nodes=100',
            '--num-classes=6', '--hidden-dim=30', '--output-dim=30', '--assign-ratio=0.1', '--num-pool=1',
            '--epochs=1', '--num_workers=0', '--cuda=0'] + %(extra)r
import train
train.log_assignment = lambda *a, **k: None
train.log_graph = lambda *a, **k: None
import cross_val
graphs_seen = []
orig = train.train
def one_fold(*a, **k):                      # benchmark_task_val runs 10 folds; one is enough here
    r = orig(*a, **k)
    graphs_seen.append(len(r[1]))
    raise SystemExit(0)
train.train = one_fold
try:
    train.main()
except SystemExit:
    pass
assert graphs_seen == [1], graphs_seen
print('SHIM_OK', sorted(stubbed))
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, 'train.py')), reason='reference checkout not present')
@pytest.mark.parametrize('method,extra', [('soft-assign', ['--linkpred']), ('base-set2set', ['--dropout=0.1'])])
def test_reference_train_py_runs_one_epoch_under_the_shim(tmp_path, method, extra):
    r = subprocess.run([sys.executable, '-c', SCRIPT % {'root': ROOT, 'ref': REF, 'method': method, 'extra': extra}],
                       cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    assert 'SHIM_OK' in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
    assert 'Validation  accuracy' in r.stdout


def test_networkx_compat_and_stubs():
    from graph_pooling_b200 import shim
    nx = shim.install_networkx_compat()
    g = nx.path_graph(3)
    g.node[0]['feat'] = 1
    assert g.nodes[0]['feat'] == 1 and float(nx.__version__) > 1.0
    assert nx.to_numpy_matrix(g).shape == (3, 3) and nx.from_numpy_matrix(nx.to_numpy_matrix(g)).number_of_edges() == 2
    o = shim._NullObj()
    assert o.figure().add_subplot(1, 2)[0].plot([1]) is o and list(o) == []
