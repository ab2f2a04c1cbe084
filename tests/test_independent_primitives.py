"""The 3GPP primitives that the synthetic transmitter (sim/) and the CPU oracle SHARE (sim/lte_common.c) against implementations that are not ours.
A mistake in one of them would be invisible to every transmitter -> oracle -> CUDA comparison, because both ends would make it; srsRAN, which would pin
them, is absent.  Third-party code / published constants used here:
  * CRC: binascii.crc_hqx (CRC-16/XMODEM = gCRC16 of 36.212 5.1.1) and the check values of the CRC catalogue for CRC-24/LTE-A, CRC-24/LTE-B, CRC-8/LTE;
  * Gold sequences (36.211 7.2): scipy.signal.max_len_seq as the LFSR;
  * OFDM demodulation: numpy.fft;
and, where no library exists, a formulation in another domain than the shift-register loops of lte_common.c (polynomial products over GF(2) with
numpy.convolve for the convolutional and turbo encoders, closed-form constellation and Zadoff-Chu expressions of 36.211 evaluated in float64)."""
import binascii
import ctypes as C
import os
import sys
import numpy as np
import pytest
import ltelib
from ltelib import Cell

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import check_tables  # noqa: E402


@pytest.fixture(scope="module")
def S(infra):
    L = infra.sim()
    L.lte_crc.restype = C.c_uint32
    L.lte_crc.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    L.lte_gold_bits.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32]
    L.lte_conv_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    L.lte_turbo_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lte_qpp.argtypes = [C.c_uint32, C.c_void_p]
    L.lte_modulate.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    L.lte_crs.restype = C.c_uint32
    L.lte_crs.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.lte_pusch_dmrs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    return L


def bits_of(data):
    return np.unpackbits(np.frombuffer(bytes(data), np.uint8))


def test_crc_polynomials_against_binascii_and_catalogue_check_values(S):
    check = bits_of(b"123456789")
    p = ltelib.ptr
    # published check values (CRC RevEng catalogue): CRC-24/LTE-A, CRC-24/LTE-B, CRC-16/XMODEM, CRC-8/LTE
    assert S.lte_crc(0x1864CFB, 24, p(check), 72) == 0xCDE703
    assert S.lte_crc(0x1800063, 24, p(check), 72) == 0x23EF52
    assert S.lte_crc(0x11021, 16, p(check), 72) == 0x31C3
    assert S.lte_crc(0x19B, 8, p(check), 72) == 0xEA
    rng = np.random.default_rng(0)
    for n in (1, 2, 7, 31, 100, 753):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        b = bits_of(data)
        assert S.lte_crc(0x11021, 16, p(b), 8 * n) == binascii.crc_hqx(data, 0)
    # bit lengths that are no multiple of 8 (DCI payloads): a CRC of the zero-extended message through the library, undone by the linearity of the code
    for nb in (27, 31, 43):
        b = rng.integers(0, 2, nb).astype(np.uint8)
        padded = np.concatenate([np.zeros((-nb) % 8, np.uint8), b])           # leading zeros do not change a zero-initialised CRC
        assert S.lte_crc(0x11021, 16, p(b), nb) == binascii.crc_hqx(np.packbits(padded).tobytes(), 0)


def gold_scipy(c_init, n):
    from scipy.signal import max_len_seq
    x1 = np.zeros(31, np.int8)
    x1[0] = 1
    x2 = np.array([(c_init >> i) & 1 for i in range(31)], np.int8)
    s1 = max_len_seq(31, state=x1, length=1600 + n, taps=[3])[0]            # x1(n+31) = x1(n+3) + x1(n)
    if c_init == 0:                                                        # the all-zero register stays zero (scipy refuses to run it)
        s2 = np.zeros(1600 + n, s1.dtype)
    else:
        s2 = max_len_seq(31, state=x2, length=1600 + n, taps=[1, 2, 3])[0]  # x2(n+31) = x2(n+3) + x2(n+2) + x2(n+1) + x2(n)
    return (s1[1600:] ^ s2[1600:]).astype(np.uint8)


def test_gold_sequence_against_scipy_lfsr(S):
    rng = np.random.default_rng(1)
    for c_init in [1, 0x7FFFFFFF, (0x1234 << 14) + (3 << 9) + 301] + [int(x) for x in rng.integers(1, 1 << 31, 6)]:
        n = 5000
        out = np.zeros(n, np.uint8)
        S.lte_gold_bits(c_init, ltelib.ptr(out), n)
        assert np.array_equal(out, gold_scipy(c_init, n)), hex(c_init)


def test_crs_values_follow_the_gold_sequence(S):
    """36.211 6.10.1.1: r(m) = ((1 - 2 c(2m)) + j (1 - 2 c(2m+1))) / sqrt 2, c_init = 2^10 (7 (ns+1) + l + 1)(2 N_id + 1) + 2 N_id + N_CP, the 2 nof_prb values
    of the cell cut from the middle of the 220 of the maximum bandwidth"""
    for cellp in ((100, 2, 301, 2), (25, 1, 5, 1), (50, 2, 167, 2)):
        cell = Cell(*cellp)
        for ns, l in ((0, 0), (7, 4), (19, 0), (12, 4)):
            pil = np.zeros(2 * cell.nof_prb, np.complex64)
            for port in range(cell.nof_ports):                            # frequency position: k = 6 m + (v + v_shift) mod 6, v = 0 / 3 by port and symbol
                off = S.lte_crs(C.byref(cell), port, ns, l, ltelib.ptr(pil))
                assert off == (((0 if l == 0 else 3) if port == 0 else (3 if l == 0 else 0)) + cell.cell_id % 6) % 6
            c_init = (1 << 10) * (7 * (ns + 1) + l + 1) * (2 * cell.cell_id + 1) + 2 * cell.cell_id + 1
            c = gold_scipy(c_init, 440).astype(np.float64)
            r = ((1 - 2 * c[0::2]) + 1j * (1 - 2 * c[1::2])) / np.sqrt(2)
            m0 = 110 - cell.nof_prb
            assert np.allclose(pil, r[m0:m0 + 2 * cell.nof_prb], atol=1e-6)


def test_tail_biting_convolutional_encoder_is_a_circular_gf2_convolution(S):
    """36.212 5.1.3.1: generators 133, 171, 165 (octal), register initialised with the last six bits = circular convolution of the block with each generator"""
    gens = [[1, 0, 1, 1, 0, 1, 1], [1, 1, 1, 1, 0, 0, 1], [1, 1, 1, 0, 1, 0, 1]]
    rng = np.random.default_rng(2)
    for K in (27, 31, 43, 40, 56):
        c = rng.integers(0, 2, K).astype(np.uint8)
        out = np.zeros(3 * K, np.uint8)
        S.lte_conv_encode(ltelib.ptr(c), K, ltelib.ptr(out))
        for i, g in enumerate(gens):
            d = sum(g[j] * np.roll(c.astype(np.int64), j) for j in range(7)) % 2
            assert np.array_equal(out[i * K:(i + 1) * K], d), (K, i)


def rsc_numpy(x):
    """constituent encoder of 36.212 5.1.3.2.1, g0 = 1 + D^2 + D^3 (feedback), g1 = 1 + D + D^3, as power-series products over GF(2):
    a = x / g0 = x * (1 / g0), z = a * g1, plus the three termination steps (input = feedback, so that the register empties)"""
    K = len(x)
    inv = np.zeros(K + 3, np.int64)          # 1 / g0: s_k = s_{k-2} + s_{k-3}, s_0 = 1 (period 7)
    for k in range(K + 3):
        inv[k] = 1 if k == 0 else (inv[k - 2] if k >= 2 else 0) ^ (inv[k - 3] if k >= 3 else 0)
    a = np.convolve(x.astype(np.int64), inv)[:K] % 2
    a_ext = np.concatenate([a, np.zeros(3, np.int64)])                      # a_K .. a_{K+2} = 0 by construction of the termination
    xt = [int(a_ext[K + i - 2] ^ a_ext[K + i - 3]) for i in range(3)]       # tail inputs = the feedback value
    z = np.convolve(a_ext, [1, 1, 0, 1])[:K + 3] % 2
    return z[:K].astype(np.uint8), xt, [int(v) for v in z[K:K + 3]]


def test_turbo_encoder_against_power_series_formulation(S):
    f1, f2, _, _ = check_tables.load()
    Ks = check_tables.K_TABLE if hasattr(check_tables, "K_TABLE") else None
    rng = np.random.default_rng(3)
    for K in (40, 104, 512, 1056, 6144, 5824):
        pi = np.zeros(K, np.uint16)
        S.lte_qpp(K, ltelib.ptr(pi))
        assert sorted(pi.tolist()) == list(range(K))                        # a permutation ...
        idx = [i for i in range(188) if (40 + 8 * i if i < 60 else 512 + 16 * (i - 59) if i < 92 else 1024 + 32 * (i - 91) if i < 124 else 2048 + 64 * (i - 123)) == K][0]
        i = np.arange(K, dtype=np.int64)
        assert np.array_equal(pi, (f1[idx] * i + f2[idx] * i * i) % K)      # ... and the quadratic one of 36.212 Table 5.1.3-3
        x = rng.integers(0, 2, K).astype(np.uint8)
        d = [np.zeros(K + 4, np.uint8) for _ in range(3)]
        S.lte_turbo_encode(ltelib.ptr(x), K, ltelib.ptr(d[0]), ltelib.ptr(d[1]), ltelib.ptr(d[2]))
        z, xt, zt = rsc_numpy(x)
        zp, xpt, zpt = rsc_numpy(x[pi])
        assert np.array_equal(d[0][:K], x) and np.array_equal(d[1][:K], z) and np.array_equal(d[2][:K], zp), K
        # trellis termination multiplexing, 36.212 5.1.3.2.2
        assert [int(v) for v in d[0][K:]] == [xt[0], zt[1], xpt[0], zpt[1]]
        assert [int(v) for v in d[1][K:]] == [zt[0], xt[2], zpt[0], xpt[2]]
        assert [int(v) for v in d[2][K:]] == [xt[1], zt[2], xpt[1], zpt[2]]


def test_modulation_mapper_against_closed_forms(S):
    """36.211 7.1.2 - 7.1.5 written as formulas instead of tables: every axis is a Gray-coded PAM, I from the even bits, Q from the odd ones"""
    def pam(bits):                                                          # bits[0] = sign, the following ones fold the amplitude
        v = np.zeros(bits.shape[1])
        for k in range(bits.shape[0] - 1, 0, -1):
            v = (1 << (bits.shape[0] - k)) - (1 - 2.0 * bits[k]) * (v if k < bits.shape[0] - 1 else 1.0)
        return (1 - 2.0 * bits[0]) * (v if bits.shape[0] > 1 else 1.0)
    rng = np.random.default_rng(4)
    for qm, norm in ((2, 2.0), (4, 10.0), (6, 42.0), (8, 170.0)):
        n = 4096
        b = rng.integers(0, 2, n * qm).astype(np.uint8)
        b[:qm * (1 << qm)] = np.array([[(s >> (qm - 1 - k)) & 1 for k in range(qm)] for s in range(1 << qm)], np.uint8).ravel()   # every point once
        out = np.zeros(n, np.complex64)
        S.lte_modulate(ltelib.ptr(b), n, qm, ltelib.ptr(out))
        bb = b.reshape(n, qm).T.astype(np.float64)
        ref = (pam(bb[0::2]) + 1j * pam(bb[1::2])) / np.sqrt(norm)
        assert np.allclose(out, ref, atol=1e-6), qm
        assert abs(np.mean(np.abs(out[:1 << qm]) ** 2) - 1.0) < 1e-5         # unit average power over the constellation


def test_pusch_dmrs_against_zadoff_chu_formula(S):
    """36.211 5.5.1.1 / 5.5.2.1.1 for M_sc >= 36 without hopping: r(n) = exp(j alpha n) x_q(n mod N_zc), x_q(m) = exp(-j pi q m (m+1) / N_zc),
    q from u = (cell_id mod 30 + delta_ss) mod 30, alpha = 2 pi ((n_dmrs1 + n_dmrs2 + n_prs(ns)) mod 12) / 12"""
    n1_map = [0, 2, 3, 4, 6, 8, 9, 10]                                      # cyclicShift -> n_DMRS^(1), Table 5.5.2.1.1-2
    for cell_id, delta_ss, cs, n2, nprb, ns in ((301, 0, 0, 0, 3, 0), (5, 7, 3, 6, 10, 11), (167, 29, 7, 9, 25, 19), (40, 3, 5, 4, 48, 4)):
        cell = Cell(50, 1, cell_id, 1)
        ucfg = (C.c_uint32 * 5)(cs, delta_ss, 0, 0, 0)
        M = 12 * nprb
        r = np.zeros(M, np.complex64)
        assert S.lte_pusch_dmrs(C.byref(cell), ucfg, ns, n2, M, ltelib.ptr(r)) == 0
        nzc = max(p for p in range(2, M) if all(p % d for d in range(2, int(p ** 0.5) + 1)))
        u = ((cell_id % 30) + delta_ss) % 30
        qbar = nzc * (u + 1) / 31.0
        q = int(np.floor(qbar + 0.5))                                        # v = 0
        m = np.arange(M) % nzc
        base = np.exp(-1j * np.pi * q * m * (m + 1) / nzc)
        c = gold_scipy((cell_id // 30) * 32 + ((cell_id % 30) + delta_ss) % 30, 8 * 7 * 20 + 8)
        nprs = sum(int(c[8 * 7 * ns + i]) << i for i in range(8))
        alpha = 2 * np.pi * ((n1_map[cs] + n2 + nprs) % 12) / 12
        assert np.allclose(r, np.exp(1j * alpha * np.arange(M)) * base, atol=2e-4), (cell_id, nprb)


def test_ofdm_demodulator_against_numpy_fft(infra):
    """K1 of the oracle: cyclic prefixes 160 / 144 (scaled), FFT of symbol_sz points, the 12 nof_prb carriers around DC (DC itself skipped), no scaling"""
    rng = np.random.default_rng(5)
    for cellp, symsz in (((25, 1, 5, 1), 0), ((50, 2, 3, 2), 0), ((100, 2, 1, 2), 0), ((100, 2, 1, 2), 1536)):
        cell = Cell(*cellp, symsz)
        o = ltelib.Oracle(cell)
        N = cell.fft()
        sf_len = 15 * N
        iq = (rng.standard_normal(sf_len) + 1j * rng.standard_normal(sf_len)).astype(np.complex64)
        sym = o.ofdm(np.stack([iq] * cell.nof_rx))[0].reshape(14, 12 * cell.nof_prb)
        nsc = 12 * cell.nof_prb
        pos = 0
        for l in range(14):
            cp = (160 if l % 7 == 0 else 144) * N // 2048
            X = np.fft.fft(iq[pos + cp:pos + cp + N].astype(np.complex128))
            ref = np.concatenate([X[N - nsc // 2:], X[1:nsc // 2 + 1]])
            assert np.allclose(sym[l], ref, atol=2e-5 * np.sqrt(N)), (cellp, symsz, l)      # fp32 butterflies against float64
            pos += cp + N
        assert pos == sf_len


P_TURBO = [0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30, 1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31]   # 36.212 Table 5.1.4-1
P_CONV = [1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31, 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30]    # Table 5.1.4-2


def subblock(y, P):
    """sub-block interleaver as the matrix of 36.212 5.1.4.1.1: NULLs (-1) in front, rows of 32, columns permuted, read column by column"""
    D = len(y)
    R = -(-D // 32)
    m = np.concatenate([np.full(32 * R - D, -1, np.int64), np.asarray(y, np.int64)]).reshape(R, 32)
    return m[:, P].T.reshape(-1), R


def test_turbo_rate_matching_against_matrix_formulation(S):
    S.lte_rm_turbo_tx.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    rng = np.random.default_rng(6)
    for K, F in ((40, 0), (104, 8), (1056, 0), (6144, 0), (5824, 24)):
        D = K + 4
        d = rng.integers(0, 2, 3 * D).astype(np.uint8)
        lab = [np.arange(s * D, (s + 1) * D, dtype=np.int64) for s in range(3)]          # interleave positions, then look the bits up
        for s in (0, 1):
            lab[s][:F] = -1                                                 # filler bits are NULL in the systematic and first parity stream
        v0, R = subblock(lab[0], P_TURBO)
        v1, _ = subblock(lab[1], P_TURBO)
        Kpi = 32 * R
        y2 = np.concatenate([np.full(Kpi - D, -1, np.int64), lab[2]])
        k = np.arange(Kpi)
        v2 = y2[(np.array(P_TURBO)[k // R] + 32 * (k % R) + 1) % Kpi]      # second parity: the same permutation shifted by one
        w = np.empty(3 * Kpi, np.int64)
        w[:Kpi] = v0
        w[Kpi::2] = v1
        w[Kpi + 1::2] = v2
        for rv in range(4):
            k0 = R * (2 * -(-3 * Kpi // (8 * R)) * rv + 2)
            ring = np.concatenate([w[k0:], w, w, w])
            ring = ring[ring >= 0]
            for E in (D // 2 * 2, 3 * D - 12, 5 * D):
                e = np.zeros(E, np.uint8)
                S.lte_rm_turbo_tx(ltelib.ptr(d), K, F, rv, ltelib.ptr(e), E)
                assert np.array_equal(e, d[ring[:E]]), (K, F, rv, E)


def test_convolutional_rate_matching_against_matrix_formulation(S):
    S.lte_rm_conv_tx.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    rng = np.random.default_rng(7)
    for K in (27 + 16, 31 + 16, 43 + 16, 40):
        d = rng.integers(0, 2, 3 * K).astype(np.uint8)
        w = np.concatenate([subblock(np.arange(s * K, (s + 1) * K), P_CONV)[0] for s in range(3)])      # 5.1.4.2.2: the three streams one after the other
        w = w[w >= 0]
        for E in (72, 144, 288, 576):
            e = np.zeros(E, np.uint8)
            S.lte_rm_conv_tx(ltelib.ptr(d), K, ltelib.ptr(e), E)
            assert np.array_equal(e, d[np.resize(w, E)]), (K, E)


def test_code_block_segmentation_against_python_restatement(S):
    """lte_cbsegm (C, shared by transmitter and oracle) against tools/check_tables.segm (Python, written from 36.212 5.1.2 for the table checks) on every
    transport block size of the table and on sizes with filler bits"""
    class Segm(C.Structure):
        _fields_ = [(n, C.c_uint32) for n in ("tbs", "C", "Kp", "Km", "Cp", "Cm", "F")]
    S.lte_cbsegm.argtypes = [C.c_void_p, C.c_uint32]
    _, _, tbs, _ = check_tables.load()
    sizes = sorted(set(int(v) for v in np.array(tbs).ravel())) + [40 - 24, 6120, 6144 - 24 + 8, 12960, 100000, 101840]
    for v in sizes:
        s = Segm()
        assert S.lte_cbsegm(C.byref(s), v) == 0, v
        Cn, Kp, Km, Cp, Cm, F = check_tables.segm(v)
        assert (s.C, s.Kp, s.Cp, s.Cm, s.F) == (Cn, Kp, Cp, Cm, F) and (s.Km == Km or Cm == 0), (v, check_tables.segm(v))


def test_pusch_dmrs_group_and_sequence_hopping_against_formulas(S):
    """36.211 5.5.1.3 / 5.5.1.4 with the Gold bits from scipy: u = (f_gh(ns) + f_ss) mod 30 with f_gh = (sum_i c(8 ns + i) 2^i) mod 30, c_init = floor(N_id / 30);
    without group hopping and from 6 PRB on v = c(ns), c_init = floor(N_id / 30) 2^5 + f_ss; q = floor(qbar + 1/2) + v (-1)^floor(2 qbar)"""
    n1_map = [0, 2, 3, 4, 6, 8, 9, 10]
    nv = 0
    for cell_id, delta_ss, cs, n2, nprb, gh, sh in ((301, 4, 2, 3, 6, 1, 0), (77, 0, 1, 0, 12, 1, 1), (150, 9, 6, 8, 8, 0, 1), (9, 29, 0, 10, 25, 0, 1), (222, 1, 7, 2, 5, 0, 1)):
        cell = Cell(50, 1, cell_id, 1)
        M = 12 * nprb
        nzc = max(p for p in range(2, M) if all(p % d for d in range(2, int(p ** 0.5) + 1)))
        fss = ((cell_id % 30) + delta_ss) % 30
        c_gh = gold_scipy(cell_id // 30, 8 * 20)
        c_ss = gold_scipy((cell_id // 30) * 32 + fss, 8 * 7 * 20 + 8)
        for ns in (0, 3, 10, 19):
            ucfg = (C.c_uint32 * 5)(cs, delta_ss, gh, sh, 0)
            r = np.zeros(M, np.complex64)
            assert S.lte_pusch_dmrs(C.byref(cell), ucfg, ns, n2, M, ltelib.ptr(r)) == 0
            fgh = sum(int(c_gh[8 * ns + i]) << i for i in range(8)) % 30 if gh else 0
            u = (fgh + fss) % 30
            v = int(c_ss[ns]) if (not gh and sh and M >= 72) else 0
            nv += v
            qbar = nzc * (u + 1) / 31.0
            q = int(np.floor(qbar + 0.5)) + v * (-1) ** int(np.floor(2 * qbar))
            m = np.arange(M) % nzc
            nprs = sum(int(c_ss[8 * 7 * ns + i]) << i for i in range(8))
            alpha = 2 * np.pi * ((n1_map[cs] + n2 + nprs) % 12) / 12
            ref = np.exp(1j * alpha * np.arange(M)) * np.exp(-1j * np.pi * q * m * (m + 1) / nzc)
            assert np.allclose(r, ref, atol=2e-4), (cell_id, nprb, gh, sh, ns)
    assert nv >= 2                                                         # the second base sequence (v = 1) was exercised


def test_pusch_channel_interleaver_against_the_procedure_of_36212(S):
    """lte_uci_map (shared by transmitter and oracle) against 36.212 5.2.2.8 written down step by step in Python: rank-indication symbols from the bottom row up
    in columns {1, 4, 7, 10} taken in the order j = 0, 3, 2, 1; CQI then data row by row around them; HARQ-ACK symbols overwriting from the bottom row up in
    columns {2, 3, 8, 9} in the same order"""
    class Layout(C.Structure):
        _fields_ = [("Qp_ack", C.c_uint32), ("Qp_ri", C.c_uint32), ("Qp_cqi", C.c_uint32), ("G", C.c_uint32)]
    S.lte_uci_map.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    for M, qa, qr, qc in ((36, 0, 0, 0), (36, 5, 0, 0), (36, 0, 7, 0), (72, 9, 6, 40), (120, 33, 17, 101), (48, 4 * 48, 4 * 48, 30), (36, 1, 1, 1)):
        R = M
        kind = np.zeros(R * 12, np.int64)
        pos = np.full(R * 12, -1, np.int64)
        for cols, q, k in (((1, 4, 7, 10), qr, 2),):
            i, j, r = 0, 0, R - 1
            while i < q:
                kind[r * 12 + cols[j]] = k
                i += 1
                r = R - 1 - i // 4
                j = (j + 3) % 4
        stream = [(1, i) for i in range(qc)] + [(0, i) for i in range(12 * R - qr - qc)]
        it = iter(stream)
        for r in range(R):
            for c in range(12):
                if kind[r * 12 + c] == 2:
                    continue
                kk, p = next(it)
                kind[r * 12 + c], pos[r * 12 + c] = kk, p
        under = kind.copy()
        i, j, r = 0, 0, R - 1
        cols = (2, 3, 8, 9)
        while i < qa:
            kind[r * 12 + cols[j]] = 3
            i += 1
            r = R - 1 - i // 4
            j = (j + 3) % 4
        L = Layout(qa, qr, qc, 0)
        k_out = np.zeros(R * 12, np.uint8)
        d_out = np.zeros(R * 12, np.uint32)
        S.lte_uci_map(M, C.byref(L), ltelib.ptr(k_out), ltelib.ptr(d_out))
        assert np.array_equal(k_out == 4, (kind == 3) & (under == 1))       # lte_uci_map labels an ACK symbol that punctures CQI instead of data with 4
        assert np.array_equal(np.where(k_out == 4, 3, k_out), kind), (M, qa, qr, qc)
        sel = (kind == 0) | (kind == 1) | ((kind == 3) & (under == 0))
        assert np.array_equal(d_out[sel].astype(np.int64), pos[sel]), (M, qa, qr, qc)


def test_turbo_decoder_waterfall_is_where_the_literature_puts_it(infra):
    """An external yardstick for the oracle's decoder (the thing every CUDA turbo result is compared with): the rate-1/3 LTE turbo code with K = 6144 on BPSK /
    AWGN reaches a block error rate of 1e-2 near Eb/N0 = 0.5 dB with log-MAP and some tenths of a dB later with max-log-MAP at 8 iterations (3GPP R1 turbo-code
    evaluations; Shannon limit of the rate: -0.5 dB).  So every block must decode at 1.0 dB and none at 0.0 dB; a broken extrinsic exchange, interleaver or
    termination moves the waterfall by far more than that."""
    S, O = infra.sim(), infra.oracle()
    rng = np.random.default_rng(11)
    tbs, qm, G = 6120, 2, 3 * (6144 + 4)
    res = {}
    for ebn0 in (0.0, 1.0):
        sigma = np.sqrt(1 / (2 * 10 ** ((ebn0 + 10 * np.log10(tbs / G)) / 10)))
        ok = 0
        for _ in range(10):
            pl = rng.integers(0, 256, tbs // 8).astype(np.uint8)
            e = np.zeros(G, np.uint8)
            assert S.lte_sim_dlsch_encode(ltelib.ptr(pl), tbs, 0, G, qm, 1, ltelib.ptr(e)) == 0
            x = 2.0 * e - 1.0 + sigma * rng.standard_normal(G)
            llr = np.clip(np.round(x * 8 / sigma ** 2), -32000, 32000).astype(np.int16)
            out = np.zeros(tbs // 8 + 8, np.uint8)
            it = np.zeros(32, np.uint32)
            r = O.lteo_dlsch_decode(ltelib.ptr(llr), G, tbs, 0, qm, 1, 8, 1, ltelib.ptr(out), ltelib.ptr(it))
            ok += int(r == 1 and np.array_equal(out[:tbs // 8], pl))
        res[ebn0] = ok
    assert res == {0.0: 0, 1.0: 10}, res


def test_tail_biting_viterbi_performance_is_where_the_literature_puts_it(infra):
    """the same yardstick for the oracle's DCI decoder (rate-dematch, 8-bit quantisation, tail-biting Viterbi over three copies, CRC16): the K = 7, rate-1/3
    tail-biting code with a 43-bit block reaches a block error rate of 1e-2 near Eb/N0 = 2 - 2.5 dB (3GPP short-block evaluations).  Measured here once:
    15 % of the blocks at -2 dB, 64 % at 0 dB, 98.7 % at 2 dB, all at 4 dB."""
    S = infra.sim()
    o = ltelib.Oracle(Cell(100, 2, 5, 1))
    rng = np.random.default_rng(2)
    nb, L = 27, 1
    E = 72 << L
    res = {}
    for ebn0 in (-2.0, 4.0):
        sigma = np.sqrt(1 / (2 * 10 ** ((ebn0 + 10 * np.log10((nb + 16) / E)) / 10)))
        ok = 0
        for _ in range(150):
            b = rng.integers(0, 2, nb).astype(np.uint8)
            rnti = int(rng.integers(1, 65535))
            e = np.zeros(E, np.uint8)
            S.lte_sim_pdcch_encode(ltelib.ptr(b), nb, rnti, L, ltelib.ptr(e))
            r, bits, crc = o.dci_decode((2.0 * e - 1.0 + sigma * rng.standard_normal(E)).astype(np.float32), nb)
            ok += int(r == 0 and crc == rnti and np.array_equal(bits, b))
        res[ebn0] = ok
    assert res[4.0] >= 149 and res[-2.0] <= 45, res


def test_uplink_ofdm_demodulator_against_numpy_fft(infra):
    """K1-UL of the oracle: the half-subcarrier shift of SC-FDMA (36.211 5.6) removed by exp(-j pi n / N) before the FFT, carriers k - N_sc / 2 without a DC gap"""
    O = infra.oracle()
    O.lteo_ul_ofdm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(8)
    for cellp, symsz in (((25, 1, 5, 1), 0), ((100, 1, 1, 1), 0), ((50, 1, 3, 1), 768)):
        cell = Cell(*cellp, symsz)
        o = ltelib.Oracle(cell)
        N, nsc = cell.fft(), 12 * cell.nof_prb
        iq = (rng.standard_normal(15 * N) + 1j * rng.standard_normal(15 * N)).astype(np.complex64)
        sym = np.zeros(14 * nsc, np.complex64)
        O.lteo_ul_ofdm(o.h, ltelib.ptr(iq), ltelib.ptr(sym))
        sym = sym.reshape(14, nsc)
        pos = 0
        n = np.arange(N)
        for l in range(14):
            cp = (160 if l % 7 == 0 else 144) * N // 2048
            X = np.fft.fft(iq[pos + cp:pos + cp + N].astype(np.complex128) * np.exp(-1j * np.pi * n / N))
            ref = X[(np.arange(nsc) + N - nsc // 2) % N]
            assert np.allclose(sym[l], ref, atol=2e-5 * np.sqrt(N)), (cellp, symsz, l)
            pos += cp + N


def test_transform_deprecoder_idft_against_numpy(infra):
    """the mixed-radix (5, 3, 4, 2) Stockham inverse DFT of the PUSCH path -- the expression tree the CUDA kernel reproduces bit for bit -- against numpy.fft.ifft"""
    O = infra.oracle()
    O.lteo_idft.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(9)
    for L in (3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 18, 20, 24, 25, 27, 30, 32, 36, 40, 45, 48, 50, 54, 60, 64, 72, 75, 80, 81, 90, 96, 100):
        M = 12 * L
        x = (rng.standard_normal(M) + 1j * rng.standard_normal(M)).astype(np.complex64)
        y = np.zeros(M, np.complex64)
        assert O.lteo_idft(M, ltelib.ptr(x), ltelib.ptr(y)) == 0
        assert np.allclose(y, np.fft.ifft(x.astype(np.complex128)) * M, atol=3e-5 * np.sqrt(M) * 3), L
    assert O.lteo_idft(12 * 7, ltelib.ptr(x), ltelib.ptr(y)) == -1
