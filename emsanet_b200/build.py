"""Build libemsanet_b200.so (all CUDA kernels + the C ABI) for sm_100a, in-tree.

    python -m emsanet_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with gpurun.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libemsanet_b200.so')
SOURCES = ['api.cu', 'conv_tc.cu', 'pointwise.cu', 'upsample.cu', 'upsample_nchw.cu', 'postproc.cu', 'loss.cu', 'optim.cu', 'preproc.cu']
# arg-max / arg-min decisions are taken on expf / sqrtf / division results: IEEE-accurate math for this file
NO_FAST_MATH = {'postproc.cu', 'loss.cu', 'optim.cu', 'preproc.cu'}
# No --split-compile: with it the SAME source came out in two different code generations from build to build (the halo
# conv kernel with 80 or 93 registers — `git log -p emsanet_b200/lib/conv_tc.ptxas.log` flips between them), and the
# 80-register one is 15-35 % slower on the dominant kernel (scripts/conv_single_ab.py).  Serial NVVM is deterministic.
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--use_fast_math', '-Xcompiler', '-fPIC', '-Xptxas', '-v']
NVCC_FLAGS += os.environ.get('EB200_NVCC_EXTRA', '').split()   # experiments, e.g. -DEB200_CONV_PROBES=1


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError('nvcc not found')


def _source_digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, '..', 'include')):
        for f in sorted(os.listdir(root)):
            if f.endswith(('.cu', '.cuh', '.h')):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, 'build.sha256')
    digest = _source_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    objs = []

    def compile_one(src):
        obj = os.path.join(LIBDIR, src.replace('.cu', '.o'))
        flags = [f for f in NVCC_FLAGS if not (src in NO_FAST_MATH and f == '--use_fast_math')]
        cmd = [nvcc, *flags, '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=4) as ex:
        for obj, log in ex.map(compile_one, srcs):
            objs.append(obj)
            if verbose:
                print(log)
            with open(obj.replace('.o', '.ptxas.log'), 'w') as f:   # tracked: a code-generation change shows in git diff
                f.write(''.join(l for l in log.splitlines(True) if 'Compile time' not in l))
    cmd = [nvcc, '-shared', '-o', LIB, *objs, '-lcudart']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(stamp, 'w') as f:
        f.write(digest)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
