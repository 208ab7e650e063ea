"""Build the sm_100a shared library in-tree:  python -m t2onet_b200.build [--force]

Plain nvcc (cross-compiles without a GPU); the resulting t2onet_b200/lib/libt2o_b200.so is a
C-ABI library (include/t2o.h) with no torch or Python dependency.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libt2o_b200.so')
SOURCES = ['t2o_chain.cu', 't2o_step.cu',
           't2o_score.cu', 't2o_nm.cu', 't2o_ssim.cu', 't2o_convert.cu', 't2o_cabi.cu']
HEADERS = ['t2o_math.cuh', 't2o_common.cuh', 't2o_nm_device.cuh', 't2o_chain_kernels.cuh', 't2o_step_kernels.cuh', os.path.join('..', '..', 'include', 't2o.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC'] + os.environ.get('T2O_NVCC_EXTRA', '').split()     # (development probes: -DT2O_RES_PROBE)


def nvcc_path():
    p = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    return p if os.path.exists(p) else None


def source_hash():
    """Content hash of everything the library is built from (mtimes do not survive a copy of the tree)."""
    import hashlib
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    for d in sorted(os.path.join(CSRC, s) for s in SOURCES + HEADERS):
        if os.path.exists(d):
            with open(d, 'rb') as f:
                h.update(os.path.basename(d).encode() + b'\0' + f.read())
    return h.hexdigest()


def is_stale():
    """The library is missing, or was built from other sources than the ones in the tree."""
    if not os.path.exists(LIB):
        return True
    try:
        with open(LIB + '.srchash') as f:
            return f.read().strip() != source_hash()
    except OSError:
        t = os.path.getmtime(LIB)             # a library built before the hash file existed: fall back to mtimes
        deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
        return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=True):
    if not force and not is_stale():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError('nvcc not found: cannot build %s' % LIB)
    os.makedirs(LIBDIR, exist_ok=True)
    # one builder at a time (torchrun starts every rank at once): the others wait for the lock and find the result
    import fcntl
    with open(os.path.join(LIBDIR, '.build.lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB
            return _build_locked(nvcc, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(nvcc, verbose):
    objdir = os.path.join(LIBDIR, 'obj')
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            print('[t2onet_b200.build]', ' '.join(cmd), flush=True)
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    tmp = LIB + '.tmp'
    cmd = [nvcc, '-shared', '-o', tmp] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    if verbose:
        print('[t2onet_b200.build]', ' '.join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)
    with open(LIB + '.srchash', 'w') as f:
        f.write(source_hash())
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
