"""Build the C-ABI CUDA library ``torch_nf_b200/_C.so`` in-tree with nvcc for sm_100a.

The library exports exactly the ``extern "C"`` entry points declared in
``include/tnf.h``; it has no torch or Python dependency (plain pointers and
sizes), so it is loaded with ``ctypes``.
"""
import fcntl
import hashlib
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "_C.so")
STAMP = os.path.join(HERE, "_C.so.stamp")
SOURCES = ["elementwise.cu", "coupling_generic.cu", "coupling_tc.cu", "coupling_tc4.cu", "coupling_tc5.cu", "coupling_tc6.cu", "coupling_tcb.cu", "chain.cu", "cde_fused.cu", "cde_fused_tc.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--cudart", "static",
]
OBJ = os.path.join(HERE, "_obj")      # per-source objects (git- and gpurun-ignored): only changed sources recompile


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def source_digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "tnf.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())   # not the absolute path: the tree is copied to other roots
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _headers_digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "tnf.h"))
    for f in files:
        with open(f, "rb") as fh:
            h.update(os.path.basename(f).encode())
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_objects(nvcc, verbose):
    """nvcc -c every source whose (source, headers, flags) digest changed, in parallel; returns the object paths."""
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ, exist_ok=True)
    hd = _headers_digest()
    jobs, objs = [], []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, src[:-3] + ".o")
        with open(path, "rb") as fh:
            dig = hashlib.sha256(fh.read() + hd.encode()).hexdigest()
        objs.append(obj)
        stamp = obj + ".stamp"
        if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
            continue
        jobs.append((path, obj, stamp, dig))

    def run(job):
        path, obj, stamp, dig = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", "-o", obj, path]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (os.path.basename(path), res.stdout, res.stderr))
        with open(stamp, "w") as fh:
            fh.write(dig)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(run, jobs))
    return objs


def is_current():
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == source_digest()


def build(force=False, verbose=False):
    """Compile every CUDA source into one shared library. Returns its path.

    Safe under concurrent callers (torchrun starts one process per GPU): an exclusive file lock serialises the
    builders, the library is written to a temporary name and renamed into place, and late comers find it current."""
    if not force and is_current():
        return LIB
    nvcc = _nvcc()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build torch_nf_b200/_C.so")
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and is_current():      # another process built it while we waited
                return LIB
            tmp = LIB + ".tmp.%d" % os.getpid()
            objs = _compile_objects(nvcc, verbose)
            cmd = [nvcc] + NVCC_FLAGS + ["-shared", "-Xlinker", "--no-undefined", "-o", tmp] + objs
            if verbose:
                print(" ".join(cmd).replace(tmp, LIB))
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc (link) failed:\n" + res.stdout + res.stderr)
            os.replace(tmp, LIB)
            with open(STAMP + ".tmp", "w") as fh:
                fh.write(source_digest())
            os.replace(STAMP + ".tmp", STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
