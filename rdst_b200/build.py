"""In-tree build of librdst_b200.so (nvcc, sm_100a only).  `python -m rdst_b200.build` or __graft_entry__.build()."""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "librdst_b200.so")

# no --use_fast_math: the fp32 (1e-4 parity) path relies on IEEE erff / expf / division
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# probe builds (timing experiments): RDST_NVCC_EXTRA="-DFOO" applies to the one source RDST_NVCC_EXTRA_FILE names
EXTRA = os.environ.get("RDST_NVCC_EXTRA", "").split()
EXTRA_FILE = os.environ.get("RDST_NVCC_EXTRA_FILE", "tc_attn2.cu")


def _nvcc():
    n = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(n):
        raise RuntimeError("nvcc not found; librdst_b200 cannot be built")
    return n


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stamp(src):
    h = hashlib.sha1()
    if EXTRA and os.path.basename(src) == EXTRA_FILE:
        h.update(" ".join(EXTRA).encode())
    for f in [src] + sorted(os.path.join(CSRC, x) for x in os.listdir(CSRC) if x.endswith((".cuh", ".h"))) + \
            [os.path.join(HERE, "..", "include", "rdst_b200.h")]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(name, verbose):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name[:-3] + ".o")
    stamp_file = obj + ".sha1"
    stamp = _stamp(src)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, ""
    cmd = [_nvcc()] + NVCC_FLAGS + (EXTRA if name == EXTRA_FILE else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in res]
    log = "".join(l for _, l in res)
    if verbose and log:
        print(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        cmd = [_nvcc(), "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
