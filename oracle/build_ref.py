"""Build the reference's OWN native extensions as a GPU-side oracle (test infrastructure only).

Sources are compiled where they lie under /root/reference (never copied into the tracked tree);
outputs go to the git-ignored ``oracle/_ref/`` (which still travels to the GPU box with gpurun).

* ``gridencoder_ref``: /root/reference/mycuda/torch_ngp_grid_encoder/{gridencoder.cu,bindings.cpp}
  compiled UNMODIFIED (same nvcc flags as mycuda/setup.py:18, arch sm_100a).
* ``common_ref``: /root/reference/mycuda/common.cu does not compile unmodified in this image
  (needs Eigen, absent; ``tensor.type()`` in AT_DISPATCH is rejected by torch 2.11).  The recipe
  writes a patched *generated* copy into oracle/_ref/common_src/ with (a) the Eigen include and
  the texture-baking code (common.cu:170-239, not on the hot path) cut, (b) ``.type()`` ->
  ``.scalar_type()`` at the two remaining AT_DISPATCH sites.  The two hot-path kernels are
  byte-for-byte the reference's.

Only tests/, bench.py's cpu_baseline leg and __graft_entry__ may use what this builds.
"""
import os, re, sys, time

REF = '/root/reference/mycuda'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
NVCC_FLAGS = ['-O3', '-std=c++17', '-U__CUDA_NO_HALF_OPERATORS__', '-U__CUDA_NO_HALF_CONVERSIONS__',
              '-U__CUDA_NO_HALF2_OPERATORS__']


def _load(name, sources, build_dir, extra_include=()):
    os.environ['TORCH_CUDA_ARCH_LIST'] = '10.0a'
    os.environ.setdefault('MAX_JOBS', str(os.cpu_count() or 8))
    from torch.utils.cpp_extension import load
    os.makedirs(build_dir, exist_ok=True)
    return load(name=name, sources=sources, extra_cuda_cflags=NVCC_FLAGS, extra_cflags=['-O3', '-std=c++17'],
                extra_include_paths=list(extra_include), build_directory=build_dir, verbose=True)


def build_gridencoder():
    d = f'{REF}/torch_ngp_grid_encoder'
    return _load('gridencoder_ref', [f'{d}/gridencoder.cu', f'{d}/bindings.cpp'], os.path.join(OUT, 'gridencoder'))


def build_common():
    src_dir = os.path.join(OUT, 'common_src')
    os.makedirs(src_dir, exist_ok=True)
    cu = open(f'{REF}/common.cu').read()
    cu = cu.replace('#include "Eigen/Dense"', '')
    # cut everything from the texture kernel's doc comment / template to EOF
    k = cu.index('__device__ Eigen::')          # first texture-baking helper (common.cu:170) .. EOF
    cu = cu[:k]
    cu = re.sub(r'(AT_DISPATCH_FLOATING_TYPES\(\s*\w+)\.type\(\)', r'\1.scalar_type()', cu)
    open(os.path.join(src_dir, 'common.cu'), 'w').write(cu)
    h = open(f'{REF}/common.h').read()
    h = '\n'.join(l for l in h.split('\n') if 'rayColorToTextureImageCUDA' not in l)
    open(os.path.join(src_dir, 'common.h'), 'w').write(h)
    b = open(f'{REF}/bindings.cpp').read()
    b = '\n'.join(l for l in b.split('\n') if 'rayColorToTextureImageCUDA' not in l)
    open(os.path.join(src_dir, 'bindings.cpp'), 'w').write(b)
    return _load('common_ref', [os.path.join(src_dir, 'common.cu'), os.path.join(src_dir, 'bindings.cpp')],
                 os.path.join(OUT, 'common'))


def stage_py():
    """The reference's Python that needs its CUDA extensions (GridEncoder under autocast, the occupied-voxel sampler)
    can only run on the GPU box, where /root/reference does not exist: stage the four files it needs into the
    git-ignored oracle/_ref/py/ (travels with gpurun, never tracked) for tests/golden/make_golden_gpu.py."""
    import shutil
    dst = os.path.join(OUT, 'py')
    os.makedirs(os.path.join(dst, 'mycuda', 'torch_ngp_grid_encoder'), exist_ok=True)
    for f in ('nerf_helpers.py', 'nerf_runner.py', 'Utils.py'):
        shutil.copy('/root/reference/' + f, os.path.join(dst, f))
    shutil.copy(f'{REF}/torch_ngp_grid_encoder/grid.py', os.path.join(dst, 'mycuda', 'torch_ngp_grid_encoder', 'grid.py'))


if __name__ == '__main__':
    if not os.path.isdir(REF):
        print('reference not mounted; nothing to build'); sys.exit(0)
    which = sys.argv[1:] or ['common', 'gridencoder']
    stage_py()
    for w in which:
        t = time.time()
        {'common': build_common, 'gridencoder': build_gridencoder}[w]()
        print(f'[build_ref] {w} OK in {time.time()-t:.0f}s')
