"""In-tree build of libb200env.so for sm_100a (nvcc cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
SOURCES = [os.path.join(_HERE, 'csrc', 'b200env.cu')]
HEADERS = [os.path.join(ROOT, 'include', 'b200env.h')]
OUTPUT = os.path.join(_HERE, 'libb200env.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include')]


def build_library(force=False, verbose=False):
    """Compile csrc/*.cu -> custom_envs_b200/libb200env.so if it is stale."""
    newest = max(os.path.getmtime(p) for p in SOURCES + HEADERS)
    if not force and os.path.exists(OUTPUT) and os.path.getmtime(OUTPUT) >= newest:
        return OUTPUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-o', OUTPUT] + SOURCES
    subprocess.run(cmd, check=True)
    return OUTPUT
