"""Per-C-ABI-call timing (CUDA events, profile.CallProfiler) of one eager train step at cfg1 shapes."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_pooling_b200 import dp, encoders, profile, synth  # noqa: E402

B = int(os.environ.get('PB', 4096))
dev = torch.device('cuda', 0)
batch = synth.make_batch('cfg1_enzymes_like', seed=0, device=dev, B=B)
cfg = batch['cfg']
torch.manual_seed(0)
model = synth.build_model(encoders, cfg).to(dev)
x, adj, nb, label = batch['x'], batch['adj'], batch['nb'], batch['label']


def step():
    model.zero_grad()
    yp = model(x, adj, nb, assign_x=x)
    loss = model.loss(yp, label, adj, nb)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile.CallProfiler() as cp:
    step()
torch.cuda.synchronize()
tot = 0.0
print('packed:', model._plan.packed)
for i, (name, (key, fl, by, note), e0, e1) in enumerate(cp.rows):
    ms = e0.elapsed_time(e1)
    tot += ms
    print('%3d %-24s %-40s %8.1f us' % (i, name, key[:40], ms * 1e3))
print('total %.3f ms over %d calls' % (tot, len(cp.rows)))
