"""TEST INFRASTRUCTURE: imports the UNMODIFIED reference (um-dsrg/RUMpy) so that tests, golden generators and
`bench.py --impl reference` can run it.  The product (rumpy_b200/) never imports this module.

Where the reference comes from: `baseline/_ref/` (installed by tools/install_reference.py, git-ignored, shipped to
the GPU box by gpurun) or, in the build container only, `/root/reference`.

The reference is pure Python on torch; in this image it needs two shims, neither of which touches its
arithmetic: `collections.Callable` (removed in Python 3.10, still imported by sr_tools/helper_functions.py:5) and stub
modules for third-party packages that are absent here and that the EDSR / RCAN path never calls (timm, matplotlib,
deepdiff, ...).  Three of the stubs need behaviour because the interface / trainer / evaluation hub call them:
`deepdiff.DeepDiff` (config comparison -> "no differences"), `torchinfo.summary` (model printout -> no-op) and
`prefetch_generator.BackgroundGenerator` (loader thread -> identity iterator)."""
from __future__ import annotations

import collections
import collections.abc
import importlib.abc
import importlib.machinery
import os
import shutil
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INSTALLED = os.path.join(ROOT, 'baseline', '_ref')
SOURCE = '/root/reference'
STUB_FILE = os.path.join(ROOT, 'integration', 'rumpy', 'SISR', 'models', 'b200', 'handlers.py')


class _StubLoader(importlib.abc.Loader):
    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__name__ = spec.name
        m.__path__ = []
        m.__spec__ = spec
        m.__loader__ = self
        if spec.name == 'deepdiff':
            m.DeepDiff = lambda *a, **k: {}
        elif spec.name == 'torchinfo':
            m.summary = lambda *a, **k: ''
        elif spec.name == 'prefetch_generator':
            m.BackgroundGenerator = lambda it, *a, **k: it
        return m

    def exec_module(self, module):
        pass


class _StubFinder(importlib.abc.MetaPathFinder):
    ROOTS = ('timm', 'matplotlib', 'deepdiff', 'colorama', 'torchinfo', 'prefetch_generator', 'skimage',
             'lpips', 'h5py', 'skvideo', 'moviepy', 'umap', 'click_config_file', 'imageio', 'aim', 'seaborn',
             'facenet_pytorch', 'mtcnn', 'keras', 'tensorflow', 'onnx', 'onnxruntime')

    def find_spec(self, name, path=None, target=None):
        if name.split('.')[0] in self.ROOTS:
            try:                                   # a package that IS installed wins over its stub
                for finder in sys.meta_path:
                    if finder is self or not hasattr(finder, 'find_spec'):
                        continue
                    if name.split('.')[0] == name and finder.find_spec(name, path, target) is not None:
                        return None
            except Exception:
                pass
            return importlib.machinery.ModuleSpec(name, _StubLoader(), is_package=True)
        return None


def reference_root(prefer_installed=True):
    """Directory that holds the reference's `rumpy/` tree, or None."""
    cands = [INSTALLED, SOURCE] if prefer_installed else [SOURCE, INSTALLED]
    for c in cands:
        if os.path.isfile(os.path.join(c, 'rumpy', 'shared_framework', 'models', '__init__.py')):
            return c
    return None


def available():
    return reference_root() is not None


_state = {'root': None}


def import_reference(root=None):
    """Puts the reference on sys.path (once per process) with the shims; returns the root used."""
    if _state['root'] is not None:
        if root is not None and os.path.realpath(root) != os.path.realpath(_state['root']):
            raise RuntimeError(f"reference already imported from {_state['root']}")
        return _state['root']
    root = root or reference_root()
    if root is None:
        raise RuntimeError('reference not found: run tools/install_reference.py in the build container')
    collections.Callable = collections.abc.Callable
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    sys.path.insert(0, root)
    _state['root'] = root
    return root


def overlay_with_b200_handlers(dst):
    """INTEGRATION.md section 1 carried out on a scratch copy: the unmodified reference tree from baseline/_ref plus
    the ONE new file a maintainer adds (integration/rumpy/SISR/models/b200/handlers.py).  Returns the overlay root
    (to be passed to `import_reference`)."""
    src = reference_root()
    if src is None:
        raise RuntimeError('reference not found')
    shutil.copytree(os.path.join(src, 'rumpy'), os.path.join(dst, 'rumpy'),
                    ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    cat = os.path.join(dst, 'rumpy', 'SISR', 'models', 'b200')
    os.makedirs(cat)
    shutil.copy(STUB_FILE, os.path.join(cat, 'handlers.py'))
    for d in ('Scratch', 'Results'):
        os.makedirs(os.path.join(dst, d), exist_ok=True)
    return dst


def eval_hub_kwargs(**over):
    """Keyword arguments of `EvalHub.__init__` (shared_framework/evaluation/standard_eval.py:32-39) with the defaults
    the `eval_sisr` click options give them (shared_framework/net_eval.py:19-101); `over` replaces entries."""
    kw = dict(hr_dir=None, lr_dir=None, data_attributes=None, batch_size=1, gallery_source='', galleries=('gallery_0.npz', 'gallery_1.npz', 'gallery_2.npz'),
              full_directory=False, use_celeba_blacklist=False, qpi_selection=(None, None), gallery_ref_images=None,
              dataset_name=None, group_select=None, image_shortlist=None, data_split=None, metadata_file=None,
              ignore_degradation_location=False, augmentation_normalization=None, id_source=None, recursive=False,
              model_and_epoch=(), gpu=False, sp_gpu=0, scale=4, results_name='delete_me', metrics=None, save_im=False,
              face_rec_profiling=False, model_only=False, model_loc=None, out_loc=None, no_image_comparison=False,
              save_raw_features=False, num_image_save=100000, save_data_model_folders=False, time_models=True,
              data_type='single-frame', num_frames=3, hr_selection=1, in_features=3, run_lpips_on_gpu=False,
              lanczos_upsample=False)
    unknown = set(over) - set(kw)
    if unknown:
        raise TypeError(f'not EvalHub arguments: {sorted(unknown)}')
    kw.update(over)
    return kw
