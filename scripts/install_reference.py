#!/usr/bin/env python
'''Install the UNMODIFIED reference (evalf/nutils) into baseline/_ref/ so that it travels to the GPU box.

    python scripts/install_reference.py [--reference /root/reference]

`pip install --no-index --no-build-isolation --target baseline/_ref /root/reference` cannot work in this image
(the reference's build backend flit_core is absent and there is no index), so the install is what that command
would have produced for a pure-Python package: the package tree src/nutils -> baseline/_ref/nutils, the examples ->
baseline/_ref/examples, plus the stand-ins for its absent dependencies (treelog, ags, stringly, appdirs: no-ops;
nutils_poly: the restatement of oracle/shims).  baseline/_ref/ is git-ignored: no reference source enters the history.
'''

import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def install(reference='/root/reference', quiet=False):
    src = os.path.join(reference, 'src', 'nutils')
    if not os.path.isdir(src):
        return False
    dst = os.path.join(ROOT, 'baseline', '_ref')
    os.makedirs(dst, exist_ok=True)
    for name, path in (('nutils', src), ('examples', os.path.join(reference, 'examples'))):
        target = os.path.join(dst, name)
        if os.path.isdir(target):
            shutil.rmtree(target)
        shutil.copytree(path, target, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    shims = os.path.join(ROOT, 'oracle', 'shims')
    for name in os.listdir(shims):
        if name == '__pycache__':
            continue
        s, t = os.path.join(shims, name), os.path.join(dst, name)
        if os.path.isdir(s):
            if os.path.isdir(t):
                shutil.rmtree(t)
            shutil.copytree(s, t, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        else:
            shutil.copy2(s, t)
    with open(os.path.join(dst, 'INSTALLED_FROM'), 'w') as f:
        f.write('{}\n'.format(reference))
    if not quiet:
        print('installed reference into', dst)
    return True


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--reference', default='/root/reference')
    a = ap.parse_args()
    sys.exit(0 if install(a.reference) else 1)
