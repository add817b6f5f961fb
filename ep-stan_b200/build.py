"""Builds ep-stan_b200/libepgpu.so (sm_100a) in-tree with nvcc.

    python ep-stan_b200/build.py [--force]

The shared library is git-ignored but travels to the GPU box with the snapshot.
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libepgpu.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
EXTRA = os.environ.get('EPG_NVCC_EXTRA', '').split()
FLAGS = EXTRA + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr']


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
    hdrs.append(os.path.join(os.path.dirname(HERE), 'include', 'epgpu.h'))
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + '.o')
        if force or _newer([src] + hdrs, obj):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([NVCC] + FLAGS + ['-c', src, '-o', obj], capture_output=True, text=True)
        with open(obj[:-2] + '.ptxas.log', 'w') as f:
            f.write(r.stderr)
        return src, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for src, r in ex.map(compile_one, jobs):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError('nvcc failed on ' + src)
            if verbose:
                sys.stderr.write(r.stderr)
    objs = [os.path.join(OBJ, s[:-3] + '.o') for s in srcs]
    if jobs or force or _newer(objs, LIB):
        r = subprocess.run([NVCC, '-shared', '-o', LIB] + objs + ['-lcudart'], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
