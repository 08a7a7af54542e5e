"""
Builds megadetector_b200/libmdb200.so (hand-written sm_100a CUDA + the C ABI of include/mdb200.h) in-tree with
nvcc.  No GPU is needed to build (nvcc cross-compiles); the .so is git-ignored but travels to the GPU box.

    python -m megadetector_b200.build [--force] [--verbose]
"""

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# Experiment variants: MDB_LIB_VARIANT=<name> builds libmdb200_<name>.so with the extra flags in MDB_NVCC_FLAGS (and
# _capi loads that file); the default build has neither.
VARIANT = os.environ.get('MDB_LIB_VARIANT', '')
OUT = os.path.join(HERE, 'libmdb200{}.so'.format('_' + VARIANT if VARIANT else ''))
OBJ_DIR = os.path.join(HERE, 'csrc', 'build' + ('_' + VARIANT if VARIANT else ''))
SOURCES = ['tmap.cu', 'conv_host.cu', 'preprocess.cu', 'aux.cu', 'nms.cu', 'jpeg.cu', 'engine.cu']
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
ARCH_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a']
CFLAGS = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden',
          '--expt-relaxed-constexpr']
if VARIANT:  # the variant's flags are remembered next to its objects, so a later load without the env rebuilds nothing
    _flags_file = os.path.join(OBJ_DIR, 'flags')
    if 'MDB_NVCC_FLAGS' in os.environ:
        os.makedirs(OBJ_DIR, exist_ok=True)
        with open(_flags_file, 'w') as _f:
            _f.write(os.environ['MDB_NVCC_FLAGS'])
    CFLAGS += open(_flags_file).read().split() if os.path.exists(_flags_file) else []


def _stamp():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ['../../include/mdb200.h']
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p) and n.endswith(('.cu', '.cuh', '.h')):
            with open(p, 'rb') as f:
                h.update(n.encode())
                h.update(f.read())
    h.update(' '.join(ARCH_FLAGS + CFLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library.  Returns the path of the .so.

    Safe against concurrent callers (torchrun ranks, producer workers): the build runs under an exclusive file lock, the
    objects go to a per-process scratch directory and the library is moved into place with an atomic rename, so no process
    can dlopen a half-written file."""
    import fcntl
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp_file = os.path.join(OBJ_DIR, 'stamp')
    stamp = _stamp()

    def fresh():
        return os.path.exists(OUT) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp

    if not force and fresh():
        return OUT
    with open(os.path.join(OBJ_DIR, 'lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():  # another process built it while we waited
                return OUT
            return _build_locked(stamp, stamp_file, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(stamp, stamp_file, verbose):
    if not os.path.exists(NVCC):
        raise RuntimeError('nvcc not found at {} and no up-to-date prebuilt libmdb200.so'.format(NVCC))
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace('.cu', '.o'))
        cmd = [NVCC] + ARCH_FLAGS + CFLAGS + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('---- {} ----\n{}\n'.format(src, out))
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed')
    tmp_out = OUT + '.tmp.{}'.format(os.getpid())
    cmd = [NVCC] + ARCH_FLAGS + ['-shared', '-cudart', 'static', '-Xcompiler', '-fPIC', '-o', tmp_out] + objs + ['-ldl', '-lpthread', '-lrt']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        if os.path.exists(tmp_out):
            os.remove(tmp_out)
        raise RuntimeError('link failed')
    os.replace(tmp_out, OUT)
    with open(stamp_file, 'w') as f:
        f.write(stamp)
    return OUT


FMT_SRC = os.path.join(CSRC, 'pyformat.c')
FMT_FLAGS = ['-O2', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', '-Wall']


def formatter_path():
    import sysconfig
    return os.path.join(HERE, '_mdbfmt' + sysconfig.get_config_var('EXT_SUFFIX'))


def build_formatter(force=False):
    """Compile csrc/pyformat.c (the host-side record formatter, a CPython extension) with gcc, in-tree.  Same locking and
    atomic-rename discipline as the CUDA library."""
    import fcntl
    import sysconfig
    out = formatter_path()
    fmt_dir = os.path.join(HERE, 'csrc', 'build')
    os.makedirs(fmt_dir, exist_ok=True)
    stamp_file = os.path.join(fmt_dir, 'fmt_stamp')
    with open(FMT_SRC, 'rb') as f:
        stamp = hashlib.sha256(f.read() + ' '.join(FMT_FLAGS).encode() + sys.version.encode()).hexdigest()

    def fresh():
        return os.path.exists(out) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp

    if not force and fresh():
        return out
    with open(os.path.join(fmt_dir, 'fmt_lock'), 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and fresh():
                return out
            tmp = out + '.tmp.{}'.format(os.getpid())
            cmd = [os.environ.get('CC', 'gcc')] + FMT_FLAGS + ['-I' + sysconfig.get_paths()['include'], FMT_SRC, '-o', tmp, '-lm']
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout)
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError('building the record formatter failed')
            os.replace(tmp, out)
            with open(stamp_file, 'w') as f:
                f.write(stamp)
            return out
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
    print(build_formatter(force='--force' in sys.argv))
