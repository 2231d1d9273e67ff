"""In-tree build of the CUDA extension: csrc/*.cu, csrc/*.cpp -> libtchem_b200.so (sm_100a).

Plain nvcc, no torch dependency: the library is a C-ABI shared object (include/tchem_b200.h).
Objects are rebuilt only when a source or header is newer.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libtchem_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _embed_math_header():
    """csrc/ic_math.cuh as a raw string literal (csrc/ic_math_embedded.inc): the specialised kernels
    that jit.cu generates #include it through NVRTC's in-memory headers."""
    src = os.path.join(CSRC, "ic_math.cuh")
    dst = os.path.join(CSRC, "ic_math_embedded.inc")
    text = open(src).read()
    body = 'R"TCHEMHDR(' + text + ')TCHEMHDR"\n'
    if not os.path.exists(dst) or open(dst).read() != body:
        with open(dst, "w") as f:
            f.write(body)


def build(verbose=False, force=False, extra_flags=()):
    os.makedirs(OBJ, exist_ok=True)
    _embed_math_header()
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cpp")))
    headers = (glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inc"))
               + glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    jobs = []
    objs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src) + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            cmd = [NVCC] + FLAGS + list(extra_flags) + ["-x", "cu", "-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        outs = list(ex.map(run, jobs))
    if verbose:
        for o in outs:
            if o.strip():
                print(o)
    if jobs or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                       "-Xcompiler", "-fPIC", "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        run(cmd)
    return LIB


if __name__ == "__main__":
    extra = [a for a in sys.argv[1:] if a not in ("-f", "-v")]
    print(build(verbose=True, force="-f" in sys.argv, extra_flags=extra))
