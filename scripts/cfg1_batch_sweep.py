"""ENZYMES-sized graphs (cfg1 shapes: N=100, D=3, H=E=30, K=10), batch sweep with the batch resident in
HBM and the whole train step (zero_grad, forward, CE + link loss, backward, clip, Adam) replayed from a CUDA graph.
Reports ms/step, graphs/s and achieved GB/s = algorithmic bytes (roofline.step_bytes: A once forward + once backward,
inputs, saved activations once each way, fp32) / time, against the measured HBM peak.

    python scripts/cfg1_batch_sweep.py [--batches 20,64,256,1024,4096,16384] > profiles/<round>_cfg1_batch_sweep.md
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graph_pooling_b200 import encoders, graphed, roofline, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', default='20,64,256,1024,4096,16384')
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--workload', default='cfg1_enzymes_like')
    args = ap.parse_args()
    dev = torch.device('cuda', 0)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hbm = 6459.0
    try:
        hbm = float(json.load(open(os.path.join(root, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    print('# %s shapes: batch sweep, fp32 schedule, CUDA-graph replay, batch resident in HBM' % args.workload)
    print()
    print('achieved GB/s = algorithmic bytes per step / time; HBM peak %.0f GB/s (MEASURED_PEAKS.json)' % hbm)
    print()
    print('| graphs/step | mean n_b | ms/step | graphs/s | us per graph | algorithmic MB/step | achieved GB/s | of HBM peak | kernels/step |')
    print('|---:|---:|---:|---:|---:|---:|---:|---:|---:|')
    for B in [int(b) for b in args.batches.split(',')]:
        batch = synth.make_batch(args.workload, seed=0, device=dev, B=B)
        cfg = batch['cfg']
        torch.manual_seed(0)
        model = synth.build_model(encoders, cfg).to(dev)
        model.precision = 0
        g = graphed.GraphedTrainStep(model, lr=1e-3, clip=2.0)
        x, adj, nb, label = batch['x'], batch['adj'], batch['nb'], batch['label']
        nbd = torch.from_numpy(np.ascontiguousarray(nb.astype(np.int32))).to(dev)
        for _ in range(5):
            g.step(x, adj, nbd, label)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        n0 = g.replayed_launches
        if os.environ.get('GP_PROFILE'):         # ncu --profile-from-start off: only the timed replays are captured
            torch.cuda.cudart().cudaProfilerStart()
        e0.record()
        for _ in range(args.steps):
            g.step(x, adj, nbd, label)
        e1.record()
        torch.cuda.synchronize()
        if os.environ.get('GP_PROFILE'):
            torch.cuda.cudart().cudaProfilerStop()
        ms = e0.elapsed_time(e1) / args.steps
        by = roofline.step_bytes(nb, cfg)
        gbs = by / (ms * 1e-3) / 1e9
        print('| %d | %.1f | %.3f | %.0f | %.2f | %.1f | %.1f | %.4f | %d |'
              % (B, float(np.mean(nb)), ms, B / (ms * 1e-3), ms * 1e3 / B, by / 1e6, gbs, gbs / hbm,
                 (g.replayed_launches - n0) // args.steps))
        del g, model, batch, x, adj
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
