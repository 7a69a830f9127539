"""Per-parameter gradient errors of the CUDA path against the fp64 oracle at the BASELINE shapes (tests/
test_gpu_baseline_shapes.py cases), next to the fp32 oracle's own errors.  Diagnostic: prints a markdown table."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_gpu_baseline_shapes as T   # noqa: E402


def main():
    cases = sys.argv[1:] or list(T.CASES)
    for name in cases:
        for prec in (0, 1):
            mc, yp, loss, res = T._candidate(name, prec)
            g64, g32 = res['f64']['grads'], res['f32']['grads']
            G = max(np.linalg.norm(v) for v in g64.values())
            print('\n### %s precision=%s  (G = %.3g, loss %.6f vs %.6f)' % (name, 'bf16' if prec else 'f32', G, loss.item(),
                                                                           res['f64']['loss']))
            print('| parameter | ||g64|| | cand rel err | cand abs err / G | fp32-oracle rel err | ok (fp32 rule) |')
            print('|---|---:|---:|---:|---:|---|')
            for k, p in mc.named_parameters():
                c = p.grad.cpu().numpy().astype(np.float64)
                sc = max(np.linalg.norm(g64[k]), 1e-30)
                a = np.linalg.norm(c - g64[k])
                eo = np.linalg.norm(g32[k].astype(np.float64) - g64[k]) / sc
                ok = a <= max(1e-5, 4 * eo) * sc + 1e-7 * G
                print('| %s | %.3g | %.3g | %.3g | %.3g | %s |' % (k, sc, a / sc, a / G, eo, 'yes' if ok else 'NO'))
            del mc, yp, loss


if __name__ == '__main__':
    main()
