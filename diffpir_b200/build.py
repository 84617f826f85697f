"""Build libdiffpir_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo snapshot)."""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdiffpir_b200.so")
STAMP = os.path.join(HERE, "libdiffpir_b200.srchash")     # sha256 of the sources the shipped .so was built from
LOCK = os.path.join(HERE, ".build.lock")
SOURCES = ["prox.cu", "unet_kernels.cu", "conv_tc.cu", "attention_tc.cu", "engine.cu", "comm.cu"]
HEADERS = ["common.cuh", "unet_kernels.cuh", "tc_ptx.cuh", os.path.join("..", "..", "include", "diffpir_b200.h")]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _source_hash():
    """Content hash of everything the library is built from (mtimes do not survive the copy to the GPU box)."""
    h = hashlib.sha256()
    for d in [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]:
        with open(d, "rb") as f:
            h.update(os.path.basename(d).encode() + b"\0" + f.read())
    return h.hexdigest()


def source_hash():
    return _source_hash()[:16]


def stamp_hash():
    try:
        return open(STAMP).read().strip()[:16]
    except OSError:
        return "<none>"


def needs_build():
    if not os.path.exists(LIB):
        return True
    try:
        return open(STAMP).read().strip() != _source_hash()
    except OSError:
        return True


def build(force=False, verbose=False):
    """Compile + link under an exclusive file lock (several ranks may import the package at once): objects go to a private
    temporary directory and the finished library is moved into place atomically, then the source hash is stamped."""
    if not force and not needs_build():
        return LIB
    import fcntl
    import tempfile
    with open(LOCK, "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another process built it while we waited
                return LIB
            flags = ["-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a",
                     "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]
            if verbose:
                flags += ["-Xptxas", "-v"]
            os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
            tmp = tempfile.mkdtemp(prefix="obj.", dir=os.path.join(HERE, "build"))
            try:
                objs, procs = [], []
                for s in SOURCES:
                    o = os.path.join(tmp, s.replace(".cu", ".o"))
                    objs.append(o)
                    cmd = [_nvcc()] + flags + ["-c", os.path.join(CSRC, s), "-o", o]
                    procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
                for s, p in procs:
                    out, _ = p.communicate()
                    if verbose or p.returncode:
                        sys.stderr.write(out)
                    if p.returncode:
                        raise RuntimeError(f"nvcc failed on {s}")
                out_lib = os.path.join(tmp, "libdiffpir_b200.so")
                cmd = [_nvcc(), "-shared", "-o", out_lib] + objs + ["-lcufft", "-ldl", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]
                r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                if r.returncode:
                    sys.stderr.write(r.stdout)
                    raise RuntimeError("link failed")
                os.replace(out_lib, LIB)
                with open(STAMP + ".tmp", "w") as f:
                    f.write(_source_hash() + "\n")
                os.replace(STAMP + ".tmp", STAMP)
            finally:
                shutil.rmtree(tmp, ignore_errors=True)
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
