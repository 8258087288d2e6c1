"""Build libp4b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m p4pdes_b200.build [--force]

The library is written to p4pdes_b200/lib/libp4b200.so (git-ignored; it travels to the GPU box
with the gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libp4b200.so")
SOURCES = ["mg.cu", "stencil.cu", "stencil_fast.cu", "transfer.cu", "vecops.cu", "fishfn.cu", "comm.cu", "mp_kernels.cu", "assembled.cu", "nk_device.cu", "bratu.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "kernels.h"), os.path.join(CSRC, "comm.h"),
           os.path.join(CSRC, "nk_solver.hpp"), os.path.join(CSRC, "ts_solver.hpp"),
           os.path.join(ROOT, "include", "p4b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        hdr_t = max(os.path.getmtime(h) for h in HEADERS)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hdr_t):
            continue
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % s)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-ldl"]
    subprocess.check_call(cmd)
    return LIB


SHIM_LIB = os.path.join(LIBDIR, "libpetsc_p4b200.so")
SHIM_SRC = os.path.join(HERE, "shim", "petscshim.c")
BINDIR = os.path.join(HERE, "bin")
REFERENCE = os.environ.get("P4B_REFERENCE", "/root/reference")
# the reference's unchanged drivers and the files each one is made of (c/ch6/makefile:5-7)
DRIVERS = {"fish": ["c/ch6/fish.c", "c/ch6/poissonfunctions.c"],
           "minimal": ["c/ch7/minimal.c", "c/ch6/poissonfunctions.c"],
           "pattern": ["c/ch5/pattern.c"],
           "heat": ["c/ch5/heat.c"]}


def build_shim(force=False):
    """libpetsc_p4b200.so: the PETSc-shaped C host layer (include/petsc.h) over libp4b200.so."""
    build(force=False)
    deps = [SHIM_SRC, os.path.join(ROOT, "include", "petsc.h"), os.path.join(ROOT, "include", "p4b200.h")]
    if (not force and os.path.exists(SHIM_LIB)
            and all(os.path.getmtime(d) < os.path.getmtime(SHIM_LIB) for d in deps)):
        return SHIM_LIB
    cmd = ["gcc", "-std=c99", "-O2", "-Wall", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), SHIM_SRC,
           "-o", SHIM_LIB, "-L", LIBDIR, "-lp4b200", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread"]
    subprocess.check_call(cmd)
    return SHIM_LIB


def build_drivers(force=False):
    """Compile the reference's UNCHANGED C drivers from where they lie under /root/reference against the shim.
    Returns {name: path}; skipped (prebuilt binaries are used) when the reference tree is absent."""
    out = {}
    build_shim(force=False)
    os.makedirs(BINDIR, exist_ok=True)
    for name, files in DRIVERS.items():
        exe = os.path.join(BINDIR, name)
        srcs = [os.path.join(REFERENCE, f) for f in files]
        if not all(os.path.exists(s) for s in srcs):
            if os.path.exists(exe):
                out[name] = exe
            continue
        if (not force and os.path.exists(exe) and os.path.getmtime(exe) > os.path.getmtime(SHIM_LIB)
                and all(os.path.getmtime(exe) > os.path.getmtime(s) for s in srcs)):
            out[name] = exe
            continue
        cmd = ["gcc", "-std=c99", "-pedantic", "-O2", "-I", os.path.join(ROOT, "include")] + srcs + [
            "-o", exe, "-L", LIBDIR, "-lpetsc_p4b200", "-lp4b200", "-Wl,-rpath,$ORIGIN/../lib", "-lm"]
        subprocess.check_call(cmd)
        out[name] = exe
    return out


EXAMPLES = {"minimal_native": "examples/minimal_native.c", "pattern_native": "examples/pattern_native.c"}


def build_examples(force=False):
    """The C hosts of examples/ (one C-ABI call per driver run) -> p4pdes_b200/bin/.  C99 -pedantic: the header is what a
    C maintainer includes."""
    build(force=False)
    os.makedirs(BINDIR, exist_ok=True)
    out = {}
    for name, rel in EXAMPLES.items():
        src, exe = os.path.join(ROOT, rel), os.path.join(BINDIR, name)
        if (force or not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src)
                or os.path.getmtime(exe) < os.path.getmtime(LIB)):
            subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-O2", "-I", os.path.join(ROOT, "include"), src,
                                   "-o", exe, "-L", LIBDIR, "-lp4b200", "-Wl,-rpath,$ORIGIN/../lib", "-lm"])
        out[name] = exe
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_shim(force="--force" in sys.argv))
    print(build_drivers(force="--force" in sys.argv))
    print(build_examples(force="--force" in sys.argv))
