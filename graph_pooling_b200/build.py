"""Builds libgp_b200.so in-tree with nvcc for sm_100a (no torch headers: the ABI is plain C)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libgp_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def sources():
    # .cu: device + C-ABI code (nvcc, sm_100a); .cpp: host-only helpers of the feed (compiled by the host compiler)
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cpp')))


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    hdrs.append(os.path.join(os.path.dirname(HERE), 'include', 'gp_b200.h'))
    srcs = sources()
    objs = [os.path.join(HERE, 'build', os.path.splitext(os.path.basename(s))[0] + '.o') for s in srcs]

    def compile_one(so):
        s, o = so
        if not force and not _stale(o, [s] + hdrs):
            return ''
        if s.endswith('.cpp'):
            cmd = [NVCC, '-O3', '-std=c++17', '-Xcompiler', '-fPIC,-pthread', '-c', s, '-o', o]
        else:
            cmd = [NVCC] + FLAGS + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (s, r.stdout, r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        logs = list(ex.map(compile_one, zip(srcs, objs)))
    if verbose:
        for l in logs:
            sys.stderr.write(l)
    with open(os.path.join(HERE, 'build', 'ptxas.log'), 'a') as f:
        f.write(''.join(logs))
    if force or _stale(LIB, objs):
        r = subprocess.run([NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                                    '-Xcompiler', '-pthread'],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n%s\n%s' % (r.stdout, r.stderr))
    return LIB


if __name__ == '__main__':
    print(build_native(force='--force' in sys.argv, verbose=True))
