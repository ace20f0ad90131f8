"""Recipe for oracle/_ref/gds_reference.zip: the reference's `gds` package, UNMODIFIED, as an importable archive --
TEST INFRASTRUCTURE.

The reference (asrvsn/gds) is pure Python: there is nothing to compile; its "build" is the package as it is, packed the
way a wheel packs it.  Run in the build container (where /root/reference is mounted) by __graft_entry__.build():

    python oracle/build_ref.py

packs /root/reference/gds/**/*.py into oracle/_ref/gds_reference.zip (oracle/_ref/ is git-ignored: reference sources
never enter this repository's history; it is not gpurun-ignored: the archive travels to the GPU box like a built .so)
and records where it came from.  oracle/ref_harness.py puts the archive on sys.path (zipimport), so bench.py can time
the reference itself on the box's host cores (`cpu_baseline.kind = "reference"`) and tests/test_golden_fresh_cpu.py can
regenerate golden vectors there.
"""
import hashlib
import os
import sys
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/gds'
DST = os.path.join(HERE, '_ref', 'gds_reference.zip')


def build_ref(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print('oracle/build_ref.py: /root/reference not mounted; keeping', DST if os.path.isfile(DST) else 'nothing')
        return os.path.isfile(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    h, n = hashlib.sha256(), 0
    with zipfile.ZipFile(DST, 'w', zipfile.ZIP_DEFLATED) as z:
        for root, dirs, files in os.walk(SRC):
            dirs.sort()
            # explicit directory entries: zipimport needs them for packages without an __init__.py (gds/ode)
            z.writestr(zipfile.ZipInfo(os.path.join('gds', os.path.relpath(root, SRC)).rstrip('./') + '/',
                                       date_time=(2020, 1, 1, 0, 0, 0)), b'')
            for f in sorted(files):
                if not f.endswith('.py'):
                    continue
                path = os.path.join(root, f)
                data = open(path, 'rb').read()
                h.update(data)
                # fixed timestamp: the archive is reproducible
                info = zipfile.ZipInfo(os.path.join('gds', os.path.relpath(path, SRC)), date_time=(2020, 1, 1, 0, 0, 0))
                info.compress_type = zipfile.ZIP_DEFLATED
                z.writestr(info, data)
                n += 1
    with open(os.path.join(HERE, '_ref', 'SOURCE'), 'w') as fh:
        fh.write(f'unmodified {SRC}/**/*.py ({n} files, sha256 of their concatenation {h.hexdigest()})\n')
    if verbose:
        print(f'oracle/build_ref.py: packed {n} files into {DST}')
    return True


if __name__ == '__main__':
    sys.exit(0 if build_ref() else 1)
