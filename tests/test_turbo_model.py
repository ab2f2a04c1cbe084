"""The arithmetic design of the GPU turbo kernel checked without a GPU: tools/turbo_model.cpp executes the kernel's own arithmetic
header (ltesniffer_b200/csrc/turbo_arith.cuh: biased state metrics, 32-bit IMAD-form adds next to packed 16x2 add-max, unsigned
compare of the double-biased LLR numerators, packed extrinsic) with bit-identical host definitions of the SIMD intrinsics, in the
kernel's window / normalisation schedule; its hard decisions after every iteration must equal the CPU oracle's (int32 max-log-MAP,
oracle/lte_oracle.c siso()) -- including saturated, adversarial inputs that drive the state metrics to the edges of the int16
range analysis in turbo_arith.cuh."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import ltelib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def model():
    so = os.path.join(ROOT, "tools", "libturbo_model.so")
    src = [os.path.join(ROOT, "tools", "turbo_model.cpp"), os.path.join(ROOT, "ltesniffer_b200", "csrc", "turbo_arith.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-o", so, src[0]], check=True)
    L = C.CDLL(so)
    L.turbo_model_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    return L


def qpp(K):
    S = ltelib.sim()
    pi = np.zeros(K, np.uint16)
    S.lte_qpp.argtypes = [C.c_uint32, C.c_void_p]
    S.lte_qpp(K, ltelib.ptr(pi))
    # recover (f1, f2) from pi(1) = f1 + f2, pi(2) = 2 f1 + 4 f2  (mod K) by search over the small f2 range of the standard
    for f2 in range(1, 1024):
        f1 = (int(pi[1]) - f2) % K
        if all((f1 * i + f2 * i * i) % K == int(pi[i]) for i in (2, 3, 5, 7, K - 1)):
            return f1, f2
    raise AssertionError("no QPP parameters for K=%d" % K)


def encode(K, rng):
    S = ltelib.sim()
    S.lte_turbo_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    b = rng.integers(0, 2, K).astype(np.uint8)
    d = [np.zeros(K + 4, np.uint8) for _ in range(3)]
    S.lte_turbo_encode(ltelib.ptr(b), K, ltelib.ptr(d[0]), ltelib.ptr(d[1]), ltelib.ptr(d[2]))
    return b, np.concatenate(d)


def oracle_bits(d, K, iters):
    O = ltelib.oracle()
    bits = np.zeros(K, np.uint8)
    ok = C.c_int(0)
    O.lteo_turbo_decode(ltelib.ptr(d), K, iters, 0, ltelib.ptr(bits), C.byref(ok))
    return bits


@pytest.mark.parametrize("K", [40, 104, 512, 1056, 2112, 6144])
@pytest.mark.parametrize("kind", ["awgn", "saturated", "adversarial"])
def test_kernel_arithmetic_model_equals_oracle(infra, model, K, kind):
    rng = np.random.default_rng(K * 7 + len(kind))
    f1, f2 = qpp(K)
    ds = []
    for cb in range(2):
        b, coded = encode(K, rng)
        tx = 1.0 - 2.0 * coded.astype(np.float64)           # bit 0 -> +1 ... the decoder's convention: positive LLR = bit 1?  sign handled below
        if kind == "awgn":
            y = -tx + rng.standard_normal(coded.shape) * 1.1   # low SNR: many marginal decisions over the iterations
            d = np.clip(np.round(y * 24), -255, 255)
        elif kind == "saturated":
            y = -tx + rng.standard_normal(coded.shape) * 0.9
            d = np.where(y > 0, 255, -255)                   # every input at the rail
        else:
            d = rng.choice([-255, 255, -255, 255, 0, 37], coded.shape)   # unrelated to any code word: extrinsics swing between the rails
        ds.append(np.ascontiguousarray(d, np.int16))
    for iters in (1, 2, 3, 8):
        want = [oracle_bits(ds[cb], K, iters) for cb in range(2)]
        got = [np.zeros(K, np.uint8), np.zeros(K, np.uint8)]
        assert model.turbo_model_decode(ltelib.ptr(ds[0]), ltelib.ptr(ds[1]), K, f1, f2, iters, ltelib.ptr(got[0]), ltelib.ptr(got[1])) == 0
        for cb in range(2):
            bad = np.nonzero(want[cb] != got[cb])[0]
            assert len(bad) == 0, "K=%d %s iters=%d code block %d: %d bits differ, first at %d" % (K, kind, iters, cb, len(bad), bad[0])
