"""Installs the UNMODIFIED reference (um-dsrg/RUMpy, /root/reference) under baseline/_ref/ (git-ignored; `gpurun`
ships it to the GPU box, where /root/reference does not exist).  Run in the build container:

    python tools/install_reference.py

Step 1 is the contract's command:
    python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref <src>
(<src> = a copy of /root/reference under /tmp: setup.py writes build/ and *.egg-info into the source tree, which is
read-only).  It succeeds, but the wheel it builds holds NO modules: setup.py asks `find_packages()` for the package
list and `rumpy/` has no `rumpy/__init__.py` (it is a namespace directory), so nothing is found -- upstream only
documents the editable install (`pip install -e .`, README.md:70), which puts the SOURCE TREE on sys.path.
Step 2 therefore materialises what the editable install exposes: the `rumpy/` tree is copied file by file next to
the dist-info, plus the Set5 example images BASELINE configs[0] names and the two directories the CLIs expect
(`Scratch/`, `Results/`, shared_framework/configuration/constants.py:4-8).  Nothing is edited; nothing under
baseline/_ref is tracked by git."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
DST = os.path.join(ROOT, 'baseline', '_ref')


def main():
    if not os.path.isdir(REF):
        raise SystemExit(f'{REF} not found: run this in the build container')
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST)
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, 'reference')
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns('build', '*.egg-info', '__pycache__', 'Data', 'GUI'))
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--find-links',
               '/opt/wheelhouse', '--target', DST, '--no-deps', src]
        print(' '.join(cmd))
        subprocess.run(cmd, check=True)
    if not os.path.isdir(os.path.join(DST, 'rumpy')):
        print('pip installed no modules (find_packages() skips the namespace directory rumpy/): copying the tree the '
              'editable install would expose')
        shutil.copytree(os.path.join(REF, 'rumpy'), os.path.join(DST, 'rumpy'),
                        ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for sub in ('hr', 'lr_random_blur'):
        shutil.copytree(os.path.join(REF, 'Data', 'example_data', 'Set5', sub),
                        os.path.join(DST, 'Data', 'example_data', 'Set5', sub))
    for d in ('Scratch', 'Results'):
        os.makedirs(os.path.join(DST, d), exist_ok=True)
    n = sum(len(f) for _, _, f in os.walk(DST))
    print(f'installed {n} files under {DST}')


if __name__ == '__main__':
    main()
