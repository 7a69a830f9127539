"""Launcher that runs the reference's own ``train.py`` (byte-unchanged) on the B200-native encoders.

    python -m graph_pooling_b200.shim --reference /path/to/graph-pooling [--precision f32|bf16] [--seed S] -- \\
           --bmname=ENZYMES --datadir=/path/to/graph-pooling/data --method=soft-assign --max-nodes=100 ...

Nothing in the reference tree is edited.  Before ``train`` is imported this module
  * makes ``import encoders`` (train.py:22) resolve to ``graph_pooling_b200.encoders``;
  * restores the networkx <= 2.3 API the reference was written against (``Graph.node``, ``to_numpy_matrix``,
    ``from_numpy_matrix``, a float-parsable ``nx.__version__`` for load_data.py:98);
  * supplies no-op stand-ins for modules the reference imports that are not installed (``matplotlib``,
    ``tensorboardX``, ``community``), and then replaces the two image-logging helpers that need a real
    matplotlib canvas (train.py:83-166) by no-ops;
  * creates ``results/`` and the log dir (train.py:255-266, :615-621), defaults ``--cuda`` to device 0
    (the reference's default '1' hides the only GPU of a one-GPU box), and optionally seeds every RNG
    (the reference seeds nothing).
"""
import argparse
import importlib
import os
import random
import sys
import types

import numpy as np


class _NullObj:
    """Absorbs any attribute access / call / iteration (plotting and logging stand-in)."""

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return self

    def __iter__(self):
        return iter(())

    def __getitem__(self, k):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _NullModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__') and name.endswith('__'):
            raise AttributeError(name)
        return _NullObj()


def _have(mod):
    try:
        importlib.import_module(mod)
        return True
    except Exception:
        return False


def install_networkx_compat():
    import networkx as nx
    if not hasattr(nx.Graph, 'node'):
        nx.Graph.node = property(lambda self: self.nodes)
    if not hasattr(nx, 'to_numpy_matrix'):
        nx.to_numpy_matrix = lambda G, *a, **k: np.asmatrix(nx.to_numpy_array(G, *a, **k))
    if not hasattr(nx, 'from_numpy_matrix'):
        nx.from_numpy_matrix = lambda A, *a, **k: nx.from_numpy_array(np.asarray(A), *a, **k)
    try:
        float(nx.__version__)
    except ValueError:                      # '3.6.1' -> '3.6' (load_data.py:98 does float(nx.__version__))
        nx.__version__ = '.'.join(nx.__version__.split('.')[:2])
    return nx


def install_stubs():
    """Returns the set of top-level modules that had to be stubbed."""
    stubbed = set()
    if not _have('matplotlib'):
        for name in ('matplotlib', 'matplotlib.colors', 'matplotlib.pyplot', 'matplotlib.backends',
                     'matplotlib.backends.backend_agg', 'matplotlib.figure', 'matplotlib.style'):
            sys.modules[name] = _NullModule(name)
        stubbed.add('matplotlib')
    if not _have('tensorboardX'):
        tb = types.ModuleType('tensorboardX')

        class SummaryWriter:                # scalars are printed by train.py anyway (train.py:226)
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, name):
                return lambda *a, **k: None

        tb.SummaryWriter = SummaryWriter
        sys.modules['tensorboardX'] = tb
        stubbed.add('tensorboardX')
    if not _have('community'):
        sys.modules['community'] = _NullModule('community')        # util.py:1 (louvain, plotting only)
        stubbed.add('community')
    return stubbed


def install(reference_dir, precision='f32', seed=None):
    """Prepare this process so that ``import train`` (the reference's) uses the CUDA encoders."""
    reference_dir = os.path.abspath(reference_dir)
    if not os.path.isfile(os.path.join(reference_dir, 'train.py')):
        raise FileNotFoundError('no train.py under %s' % reference_dir)
    if reference_dir not in sys.path:
        sys.path.insert(0, reference_dir)
    install_networkx_compat()
    stubbed = install_stubs()
    from . import encoders
    encoders.DEFAULT_PRECISION = {'f32': 0, 'bf16': 1}[precision]
    sys.modules['encoders'] = encoders
    if seed is not None:
        import torch
        random.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
    return stubbed


def run_train(reference_dir, train_args, precision='f32', seed=None, workdir=None):
    """Import the reference's train.py and run its main() with ``train_args`` (list of CLI strings)."""
    stubbed = install(reference_dir, precision, seed)
    if workdir is not None:
        os.makedirs(workdir, exist_ok=True)
        os.chdir(workdir)
    os.makedirs('results', exist_ok=True)
    if not any(a.startswith('--cuda') for a in train_args):
        train_args = list(train_args) + ['--cuda=0']
    sys.argv = ['train.py'] + list(train_args)
    train = importlib.import_module('train')
    if 'matplotlib' in stubbed:             # these two draw on a real canvas (train.py:83-166)
        train.log_assignment = lambda *a, **k: None
        train.log_graph = lambda *a, **k: None
    train.main()
    return train


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('--reference', required=True, help='checkout of JiaxuanYou/graph-pooling')
    ap.add_argument('--precision', default='f32', choices=['f32', 'bf16'])
    ap.add_argument('--seed', type=int, default=None)
    ap.add_argument('--workdir', default=None, help='where results/ and log/ are written (default: cwd)')
    ap.add_argument('train_args', nargs=argparse.REMAINDER, help='-- followed by train.py arguments')
    a = ap.parse_args()
    rest = a.train_args[1:] if a.train_args[:1] == ['--'] else a.train_args
    run_train(a.reference, rest, a.precision, a.seed, a.workdir)


if __name__ == '__main__':
    main()
