"""Builds libdrt.so (hand-written sm_100a kernels + host interpreter + C ABI) in-tree with nvcc.

-fmad=false / -ffp-contract=off: the reference is strict-IEEE Java; hit decisions must carry its bits.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdrt.so")
SOURCES = ["render.cu", "drt_api.cpp", "host_scene.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2", "-shared", "--expt-relaxed-constexpr", "--extended-lambda"]


LINK_FLAGS = ["-lz", "-ldl"]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if os.path.getmtime(os.path.join(root, f)) > t:
                return True
    return os.path.getmtime(__file__) > t


def build(force=False, verbose=False, out=None, defines=()):
    if out is None and not force and not needs_build():
        kernel_hash()
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out or LIB] + LINK_FLAGS
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libdrt.so")
    if verbose:
        print(r.stderr)
    if out is None:
        kernel_hash(refresh=True)
    return out or LIB


HASH_FILE = os.path.join(HERE, "libdrt.sass.sha1")


def kernel_hash(refresh=False):
    """sha1 of the SASS of every kernel in libdrt.so: the identity of the MACHINE CODE the instruction model in profiles/sass/ was derived
    from (host-side edits do not change it; any change of a kernel does)."""
    import hashlib
    if not refresh and os.path.exists(HASH_FILE) and os.path.getmtime(HASH_FILE) >= os.path.getmtime(LIB):
        return open(HASH_FILE).read().strip()
    cuobjdump = os.path.join(os.path.dirname(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")), "cuobjdump")
    r = subprocess.run([cuobjdump, "-sass", LIB], capture_output=True)
    if r.returncode != 0:
        return "unknown"
    # instruction and function-name lines only: the dump also carries the absolute path of the source file ("identifier = ..."), which must not matter
    body = b"\n".join(l for l in r.stdout.split(b"\n") if l.lstrip().startswith((b"/*", b"Function :")))
    h = hashlib.sha1(body).hexdigest()[:16]
    open(HASH_FILE, "w").write(h + "\n")
    return h


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
