"""Build the sm_100a shared library in-tree:  python surrogate-models_b200/build.py

nvcc cross-compiles without a GPU.  The .so lands in surrogate-models_b200/lib/ (git-ignored, but it
travels to the GPU box with the gpurun snapshot).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libsurrogate_b200.so')
SOURCES = ['capi.cu', 'krig_nll.cu', 'krig_nll_small.cu', 'krig_nll_small8.cu', 'krig_nll_n64.cu', 'krig_nll_large.cu', 'krig_nll_flow.cu', 'krig_nll_cta.cu', 'krig_fit.cu', 'krig_rescue.cu', 'krig_predict.cu', 'rbf.cu']
NVCC_FLAGS = ([f'-D{d}' for d in os.environ.get('SMB_DEFINES', '').split(',') if d]) + ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC']


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, '..', 'include', 'surrogate_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    # the register-resident likelihood kernel is statically unrolled per matrix size: one translation unit per NT = npad / 8
    units = [(src, [], src.replace('.cu', '.o')) for src in SOURCES]
    units += [('krig_nll_regs.cu', [f'-DSMB_REGS_NT={nt}'], f'krig_nll_regs_{nt}.o') for nt in (10, 9, 8, 7, 6, 5)]
    units.sort(key=lambda u: 0 if u[0] == 'krig_nll_regs.cu' else 1)  # the slowest units first
    for src, defs, objname in units:
        obj = os.path.join(LIBDIR, objname)
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *defs, '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas=-v')
        procs.append((objname, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
        if verbose:
            print(out)
    cmd = [nvcc, '-shared', '-o', LIB, *objs, '-lcudart']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
