"""Builds libltephy_b200.so (hand-written sm_100a CUDA + C++ host code) in-tree with nvcc.
Used by __graft_entry__.build() and by the tests' session fixture.  No JIT cache: the .so sits next
to this file so that it travels to the GPU box with the repo snapshot."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libltephy_b200.so")
SOURCES = ["k_frontend.cu", "k_viterbi.cu", "k_pdsch.cu", "k_turbo.cu", "k_pusch.cu", "ltephy_capi.cu", "lte_host.cpp", "host_search.cpp", "sinks.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-Wno-unused-function", "--shared", "-Xptxas", "-v"]


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libltephy_b200 cannot be built (there is no CPU fallback)")
    return p


COMPAT_OUT = os.path.join(HERE, "libltephy_srsran_compat.so")


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", f) for f in os.listdir(os.path.join(HERE, "..", "include"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        if not os.path.exists(COMPAT_OUT):
            build_compat()
        return OUT
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", OUT] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + log[-6000:])
    if verbose:
        print(log)
    build_compat()
    return OUT


def build_compat():
    """tier-2 shim (srsRAN / FALCON names over the tier-1 C-ABI): host code only, links libltephy_b200.so"""
    import shutil as _sh
    cxx = _sh.which("g++") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", COMPAT_OUT, os.path.join(CSRC, "srsran_compat.cpp"), "-L" + HERE, "-lltephy_b200",
           "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the srsRAN compatibility shim failed:\n" + (r.stdout + r.stderr)[-4000:])
    return COMPAT_OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
