"""Real-data throughput: DiffPool on the bundled ENZYMES graphs (tests/golden/dataset_enzymes.npz, produced through
the reference's loader) with the GPU-resident batch feed (data.GraphSet) and the CUDA-graph step (graphed.py).
BASELINE.json configs[0]: batch 20, hidden/output 30, assign-ratio 0.1, num_pool 1, link prediction on.

    python scripts/train_enzymes_fast.py [--epochs 20] [--precision f32|bf16] [--batch 20]

Prints per-epoch wall time (all training steps of the epoch incl. the device-side batch assembly), graphs/s and the
train / validation accuracy.  For comparison the reference's unchanged train.py through shim.py needs 0.20-0.30 s per
epoch of 27 steps on the same GPU (profiles/r1_shim_train_enzymes.md): its host-side feed dominates."""
import argparse, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graph_pooling_b200 import encoders, graphed
from graph_pooling_b200.data import GraphSet

ap = argparse.ArgumentParser()
ap.add_argument('--epochs', type=int, default=20)
ap.add_argument('--batch', type=int, default=20)
ap.add_argument('--precision', default='f32', choices=['f32', 'bf16'])
a = ap.parse_args()
z = np.load(os.path.join(ROOT, 'tests', 'golden', 'dataset_enzymes.npz'))
gs = GraphSet(z['n'], z['glabel'].astype(np.int64) - int(z['glabel'].min()), z['nlabel'], z['eptr'], z['edges'],
              int(z['num_node_labels'])).to('cuda')
torch.manual_seed(0)
rs = np.random.RandomState(0)
perm = rs.permutation(len(gs))
tr, va = perm[:540], perm[540:]
model = encoders.SoftPoolingGcnEncoder(100, 3, 30, 30, 6, 3, 30, assign_ratio=0.1).cuda()
model.precision = 1 if a.precision == 'bf16' else 0
step = graphed.GraphedTrainStep(model, lr=1e-3, clip=2.0)
adt = torch.uint8 if a.precision == 'bf16' else torch.float32


def accuracy(ids):
    c = 0
    with torch.no_grad():
        for i in range(0, len(ids), a.batch):
            x, adj, nb, lab = gs.batch(ids[i:i + a.batch], 100, adj_dtype=adt)
            c += int((model(x, adj, nb, assign_x=x).argmax(1) == lab).sum().item())
    return c / len(ids)


for ep in range(a.epochs):
    order = tr[rs.permutation(len(tr))]
    order = order[:len(order) // a.batch * a.batch]          # full batches only: one captured graph shape
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tot = torch.zeros((), device='cuda')
    for i in range(0, len(order), a.batch):
        x, adj, nb, lab = gs.batch(order[i:i + a.batch], 100, adj_dtype=adt)
        _, loss = step.step(x, adj, nb, lab)
        tot += loss
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    nsteps = len(order) // a.batch
    if ep % 5 == 0 or ep == a.epochs - 1:
        print('epoch %3d  loss %.4f  %.1f ms/epoch  %.3f ms/step  %.0f graphs/s  train acc %.3f  val acc %.3f'
              % (ep, tot.item() / nsteps, dt * 1e3, dt * 1e3 / nsteps, len(order) / dt, accuracy(tr), accuracy(va)),
              flush=True)
    else:
        print('epoch %3d  loss %.4f  %.1f ms/epoch  %.0f graphs/s' % (ep, tot.item() / nsteps, dt * 1e3, len(order) / dt),
              flush=True)
