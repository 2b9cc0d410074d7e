"""Import the reference's own Python modules (nerf_helpers.py, nerf_runner.py, Utils.py) from
/root/reference under stub modules for the third-party packages this image lacks.

TEST INFRASTRUCTURE ONLY: used by make_golden_cpu.py / make_golden_gpu.py to produce the committed
fixtures in tests/golden/. /root/reference does not exist on the GPU box, so the gpu script copies the
three .py files it needs into the git-ignored oracle/_ref/py/ first (never into the tracked tree).

Stubs: matplotlib, imageio, trimesh, open3d, transformations, ruamel.yaml, skimage  -> inert MagicMock
modules (never called on the code paths we execute); pytorch3d.transforms.se3_exp_map -> the oracle's
restatement (third-party, unpinned — see oracle/nof_oracle.py header); kaolin -> absent (the reference
itself wraps that import in try/except, Utils.py:24-27).
"""
import os
import sys
import types
from unittest import mock

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))


def _stub(name):
    m = mock.MagicMock(name=name)
    m.__name__ = name
    m.__path__ = []
    m.__all__ = []          # so `from transformations import *` imports nothing
    m.__spec__ = None
    return m


def import_reference(ref_dir=None, mycuda_common=None, mycuda_gridencoder=None):
    """Returns (nerf_helpers, nerf_runner, Utils) modules of the reference."""
    if ref_dir is None:
        ref_dir = '/root/reference' if os.path.isdir('/root/reference') else os.path.join(REPO, 'oracle', '_ref', 'py')
    sys.path.insert(0, REPO)
    from oracle import nof_oracle
    stubbed = []
    for name in ['matplotlib', 'matplotlib.pyplot', 'imageio', 'trimesh', 'open3d', 'transformations',
                 'ruamel', 'ruamel.yaml', 'skimage', 'skimage.measure']:
        if name not in sys.modules:
            sys.modules[name] = _stub(name)
            stubbed.append(name)
    p3 = types.ModuleType('pytorch3d')
    p3t = types.ModuleType('pytorch3d.transforms')
    p3t.se3_exp_map = nof_oracle.se3_exp_map
    p3t.so3_exp_map = None
    p3t.so3_log_map = None
    p3.transforms = p3t
    sys.modules['pytorch3d'] = p3
    sys.modules['pytorch3d.transforms'] = p3t
    # the reference imports its extensions as `mycuda.common` / top-level `gridencoder`
    if mycuda_common is not None or mycuda_gridencoder is not None:
        pkg = types.ModuleType('mycuda')
        pkg.__path__ = []
        if mycuda_common is not None:
            pkg.common = mycuda_common
            sys.modules['mycuda.common'] = mycuda_common
        sys.modules['mycuda'] = pkg
        if mycuda_gridencoder is not None:
            sys.modules['gridencoder'] = mycuda_gridencoder
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    import importlib
    if mycuda_common is not None or mycuda_gridencoder is not None:
        # an earlier import of the reference in this process WITHOUT its extensions (the CPU-side golden tests) left modules whose
        # `from mycuda import common` failed silently: import them afresh now that the extensions are there
        for name in ('Utils', 'nerf_helpers', 'nerf_runner'):
            sys.modules.pop(name, None)
    Utils = importlib.import_module('Utils')
    nerf_helpers = importlib.import_module('nerf_helpers')
    nerf_runner = importlib.import_module('nerf_runner')
    # the reference modules hold their own references to the stubs; take them out of sys.modules again so that product code run later in
    # the same process (`try: import trimesh` in NerfRunner.extract_mesh) does not pick up a MagicMock
    for name in stubbed:
        sys.modules.pop(name, None)
    return nerf_helpers, nerf_runner, Utils
