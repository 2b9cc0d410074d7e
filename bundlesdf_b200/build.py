"""Build libnof_sm100.so (hand-written sm_100a CUDA behind the C ABI of include/nof.h) IN-TREE.

    python -m bundlesdf_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT_DIR = os.environ.get('NOF_BUILD_DIR', os.path.join(HERE, 'lib'))     # NOF_BUILD_DIR / NOF_EXTRA_FLAGS: tuning variants
LIB = os.path.join(OUT_DIR, 'libnof_sm100.so')
SOURCES = ['nof_api.cu', 'nof_grid.cu', 'nof_sampling.cu', 'nof_pose.cu', 'nof_adam.cu', 'nof_step_amp.cu', 'nof_step_tc.cu', 'nof_step_ws.cu', 'nof_step_f32.cu', 'nof_mesh.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr', '-Xptxas', '-v'] + os.environ.get('NOF_EXTRA_FLAGS', '').split()


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh')]
    hdrs.append(os.path.join(os.path.dirname(HERE), 'include', 'nof.h'))
    return hdrs


def _stale(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _compile(src, force, verbose):
    obj = os.path.join(OUT_DIR, src.replace('.cu', '.o'))
    path = os.path.join(CSRC, src)
    if not force and not _stale(obj, [path] + _deps()):
        return obj, ''
    cmd = [NVCC] + FLAGS + ['-c', path, '-o', obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{log}')
    with open(obj + '.ptxas.log', 'w') as f:
        f.write(log)
    if verbose:
        print(log)
    return obj, log


def build(force=False, verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = [o for o, _ in ex.map(lambda s: _compile(s, force, verbose), SOURCES)]
    if force or _stale(LIB, objs):
        cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('link failed:\n' + r.stdout + r.stderr)
    return LIB


if __name__ == '__main__':
    lib = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print('built', lib)
