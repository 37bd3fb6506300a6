"""Build the C-ABI CUDA library in-tree with nvcc for sm_100a (no torch extension machinery: the boundary is a
plain shared object, see include/gpitch_b200.h)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libgpitch_b200.so')
SOURCES = ['api.cu', 'composite.cu', 'gemm.cu', 'gemm_tma.cu', 'builder.cu', 'grad_lag.cu', 'chol.cu', 'ops.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(HERE), 'include', 'gpitch_b200.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    # one nvcc -c per translation unit, all at once (no relocatable device code: the units only share host symbols), then link
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    objdir = tempfile.mkdtemp(prefix='gpx_build_')
    compile_flags = [f for f in NVCC_FLAGS if f != '-shared']

    def compile_one(src):
        obj = os.path.join(objdir, src.replace('.cu', '.o'))
        cmd = [nvcc] + compile_flags + (['-Xptxas', '-v'] if verbose else []) + ['-c', '-o', obj, os.path.join(CSRC, src)]
        return obj, subprocess.run(cmd, capture_output=True, text=True)
    try:
        with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
            results = list(ex.map(compile_one, SOURCES))
        for obj, r in results:
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError('nvcc failed compiling %s' % obj)
            if verbose:
                sys.stderr.write(r.stderr)
        r = subprocess.run([nvcc, '-shared', '-Xcompiler', '-fPIC', '-o', LIB] + [o for o, _ in results],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError('nvcc failed linking %s' % LIB)
    finally:
        import shutil
        shutil.rmtree(objdir, ignore_errors=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
