"""Builds libwgs_b200.so in-tree with nvcc for sm_100a (no torch extension machinery: the C ABI has
no torch types in it, so a plain shared library is all that is needed)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, 'libwgs_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v', '-I', os.path.join(os.path.dirname(PKG), 'include')]


def sources():
    return sorted(glob.glob(os.path.join(HERE, '*.cu')))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = (sources() + glob.glob(os.path.join(HERE, '*.cuh')) + [os.path.abspath(__file__)] +
            glob.glob(os.path.join(os.path.dirname(PKG), 'include', '*.h')))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(HERE, 'build', os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and all(os.path.getmtime(obj) > os.path.getmtime(h) for h in glob.glob(os.path.join(HERE, '*.cuh')) +
                        glob.glob(os.path.join(os.path.dirname(PKG), 'include', '*.h')))):
            continue
        cmd = [NVCC] + FLAGS + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s' % src)
        log = os.path.join(HERE, 'build', os.path.basename(src)[:-3] + '.ptxas.log')
        with open(log, 'w') as f:
            f.write(out)
    cmd = [NVCC, '-shared', '-o', LIB] + objs + ['-lcudart', '-ldl']
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
