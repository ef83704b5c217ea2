"""In-tree build of libgenewalk_b200.so with nvcc for sm_100a (no JIT cache, no fallback)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO_PATH = os.path.join(HERE, os.environ.get('GW_SO_NAME', 'libgenewalk_b200.so'))
SOURCES = ['gw_scan_sort.cu', 'gw_walk.cu', 'gw_sgns.cu', 'gw_sgns_train.cu', 'gw_graph.cu', 'gw_stats.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-fmad=false',  # every fp op rounds on its own: bit-exact against the oracle
              '-Xcompiler', '-fPIC', '--shared',
              '-cudart', 'shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError('nvcc not found')


def _deps():
    deps = [os.path.join(CSRC, f) for f in SOURCES + ['gw_common.cuh']]
    deps.append(os.path.join(HERE, '..', 'include', 'genewalk_b200.h'))
    return deps


def source_hash():
    """Content hash of everything the library is built from (mtimes do not survive a copy of the
    tree to another box; contents do)."""
    import hashlib
    h = hashlib.sha256(' '.join(NVCC_FLAGS).encode())
    for d in _deps():
        with open(d, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(SO_PATH) or not os.path.exists(SO_PATH + '.srchash'):
        return True
    with open(SO_PATH + '.srchash') as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get('GW_EXTRA_NVCC_FLAGS', '').split() + \
          (['-Xptxas', '-v'] if verbose else []) + \
          [os.path.join(CSRC, f) for f in SOURCES] + ['-o', SO_PATH]
    env = dict(os.environ)
    env.setdefault('NVCC_CCBIN', '/usr/bin/g++')
    cmd[1:1] = ['-ccbin', '/usr/bin/g++'] if os.path.exists('/usr/bin/g++') else []
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n%s\n%s' % (res.stdout, res.stderr))
    if verbose:
        print(res.stderr)
    with open(SO_PATH + '.srchash', 'w') as f:
        f.write(source_hash())
    return SO_PATH


if __name__ == '__main__':
    import sys
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
