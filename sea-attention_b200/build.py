"""Builds libsea_b200.so (hand-written sm_100a CUDA behind the C ABI of include/sea_b200.h) IN-TREE.

    python sea-attention_b200/build.py [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  cudart is linked statically and the driver API
(cuTensorMapEncodeTiled for the TMA descriptors) is resolved at run time through
cudaGetDriverEntryPoint, so the library loads on a box without libcuda (symbol-export test).
"""
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
LIB_DIR = os.path.join(PKG_DIR, 'lib')
LIB_PATH = os.path.join(LIB_DIR, 'libsea_b200.so')
OBJ_DIR = os.path.join(LIB_DIR, 'obj')
INCLUDE = os.path.join(os.path.dirname(PKG_DIR), 'include')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr',
    '-I', INCLUDE,
]


def _nvcc():
    return os.environ.get('NVCC') or shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ['../../include/sea_b200.h']:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(LIB_DIR, 'build.stamp')
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB_PATH
    extra = ['-Xptxas', '-v'] if verbose else []

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + '.o')
        cmd = [_nvcc()] + NVCC_FLAGS + extra + ['-c', src, '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [_nvcc(), '-shared', '-o', LIB_PATH] + objs + ['-cudart', 'static', '-Xlinker', '--no-undefined', '-ldl', '-lrt', '-lpthread']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    with open(stamp, 'w') as f:
        f.write(digest)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
