"""Lane-level model of the rotating butterfly layout of dci_viterbi_kernel (k_viterbi.cu), checked against the CPU oracle without a GPU.
The 64 trellis states live two per lane; instead of routing the two new states of a lane back to fixed owners (two shuffles and
four selects per step), the layout rotates: before step t (phase f = t mod 5) lane bit i holds state bit ((i + f) mod 5) + 1 and
the register index holds state bit 0; ONE exchange with lane ^ (16 >> f) brings the butterfly partners (state bit 5) together, and
the two new states of the butterfly are already in place for phase f + 1.  Path metrics are kept non-negative (branch metrics
offset by +765, renormalisation every 5 steps to state 0 + 9180) so that their adds can run as 32-bit IMADs on packed pairs.
This file restates that scheme with numpy arrays of 32 lanes -- ballot words, traceback in rotated coordinates, non-negative
metrics -- and compares decoded bits and CRC remainders with lteo_dci_decode."""
import numpy as np
import pytest
import ltelib
from ltelib import Oracle, Cell

OFFS = 765
SPREAD = 9180


def conv_outputs(p5):
    """outputs (o0, o1, o2) of the encoder for old state p (bit 5 = 0) and input 0; state bit j-1 = c_{k-j} (oracle: T(j))"""
    T = lambda j: (p5 >> (j - 1)) & 1
    return T(2) ^ T(3) ^ T(5) ^ T(6), T(1) ^ T(2) ^ T(3) ^ T(6), T(1) ^ T(2) ^ T(4) ^ T(6)


def lane_state_bits(lane, f):
    """phase f, AFTER the exchange: the 5 low state bits b4..b0 of the butterfly this lane computes (b5 is the register index)"""
    b = [0] * 6
    b[0] = (lane >> ((4 - f) % 5)) & 1
    for k in range(1, 5):
        b[k] = (lane >> ((k - 1 - f) % 5)) & 1
    return sum(b[k] << k for k in range(5))


def model_decode(r, K, nb):
    """r: int array [3][K] of quantised symbols 2q - 255 -> (bits[nb], crc_rem) with the kernel's layout and arithmetic"""
    lanes = np.arange(32)
    # per-phase pattern index of the lane's butterfly
    pat = np.zeros((5, 32), int)
    for f in range(5):
        for l in range(32):
            o = conv_outputs(lane_state_bits(l, f))
            pat[f, l] = o[0] * 4 + o[1] * 2 + o[2]
    # non-negative branch metric table S[k][pattern] = 765 + sum_i (o_i ? r_i : -r_i)
    S = np.zeros((K, 8), int)
    for k in range(K):
        for p in range(8):
            S[k, p] = OFFS + sum((r[i][k] if (p >> (2 - i)) & 1 else -r[i][k]) for i in range(3))
    assert S.min() >= 0 and S.max() <= 2 * OFFS
    a0 = np.zeros(32, int)
    a1 = np.zeros(32, int)
    dec = np.zeros((3 * K, 2), np.uint64)          # ballot words for new state bit 0 = 0 / 1
    for t in range(3 * K):
        f, k = t % 5, t % K
        d = 16 >> f
        bit = (lanes & d) != 0
        snd = np.where(bit, a0, a1)
        rcv = snd[lanes ^ d]
        na0 = np.where(bit, rcv, a0)
        na1 = np.where(bit, a1, rcv)
        m = S[k, pat[f]]
        mn = 2 * OFFS - m
        t0, t1, t2, t3 = na0 + m, na1 + mn, na0 + mn, na1 + m
        # state p (register 0 side) has b5 = 0; input u: branch metric of (p, u = 0) is m, of (p + 32, 0) is 2*765 - m
        # new state with u = 0: max(p: a0 + bm(p,0), p+32: a1 + bm(p+32,0)); the oracle's bm for input c uses sgn[p][c]: flipping c
        # flips all three outputs
        n0 = np.maximum(t0, t1)
        n1 = np.maximum(t2, t3)
        w0 = sum(int(t1[l] > t0[l]) << l for l in range(32))     # ties keep the b5 = 0 predecessor
        w1 = sum(int(t3[l] > t2[l]) << l for l in range(32))
        dec[t, 0], dec[t, 1] = w0, w1
        a0, a1 = n0, n1
        if t % 5 == 4:                     # once per group of five phases
            ref = a0[0] - SPREAD           # state 0 is register 0 of lane 0 in every phase
            a0, a1 = a0 - ref, a1 - ref
            assert a0.min() >= 0 and a1.min() >= 0 and max(a0.max(), a1.max()) <= 2 * SPREAD
        assert max(a0.max(), a1.max()) < 32768
    # best end state: layout of phase psi = (3K) mod 5 before an exchange: lane bit i = state bit ((i + psi) mod 5) + 1, register = b0
    psi = (3 * K) % 5
    best_v, best_s = -1, 0
    for l in range(32):
        for rr in range(2):
            s = rr
            for i in range(5):
                s |= ((l >> i) & 1) << (((i + psi) % 5) + 1)
            v = (a0 if rr == 0 else a1)[l]
            if v > best_v or (v == best_v and s < best_s):
                best_v, best_s = v, s
    # traceback in rotated coordinates: (y, u) = (lane, register) of the state in the layout the decisions of step t were balloted
    # in (phase (t + 1) mod 5); going back one step replaces ONE bit of y -- the one that held state bit 1 -- by the decision
    x = best_s >> 1
    y = ((x >> psi) | (x << (5 - psi))) & 31
    u = best_s & 1
    pos = (5 - psi) % 5
    st = best_s
    data = np.zeros(K, np.uint8)
    for t in range(3 * K - 1, K - 1, -1):
        psi_t = (t + 1) % 5
        xx = st >> 1
        assert y == ((xx >> psi_t) | (xx << (5 - psi_t))) & 31 and u == (st & 1) and pos == (5 - psi_t) % 5   # incremental == direct
        if t < 2 * K:
            data[t - K] = u
        dbit = (int(dec[t, u]) >> y) & 1
        st = (st >> 1) | (dbit << 5)
        u = (y >> pos) & 1
        y = (y & ~(1 << pos)) | (dbit << pos)
        pos = (pos + 1) % 5
    reg = 0
    for i in range(nb + 16):
        reg = (reg << 1) | (int(data[i]) if i < nb else 0)
        if reg & 0x10000:
            reg ^= 0x11021
    par = 0
    for i in range(nb, nb + 16):
        par = (par << 1) | int(data[i])
    return data[:nb], (par ^ reg) & 0xFFFF


@pytest.mark.parametrize("nb", [15, 28, 39, 48, 51])
def test_rotating_butterfly_layout_matches_oracle(infra, nb):
    o = Oracle(Cell(100, 2, 1, 2))
    rng = np.random.default_rng(nb)
    S = infra.sim()
    K = nb + 16
    for trial in range(6):
        L = int(rng.integers(0, 4))
        E = 72 << L
        if trial < 3:
            e = rng.standard_normal(E).astype(np.float32)                       # pure noise: every decision is marginal
        else:
            e = (np.sign(rng.standard_normal(E)) * 1.0 + 0.6 * rng.standard_normal(E)).astype(np.float32)
        rc, bits, crc = o.dci_decode(e, nb)
        assert rc == 0
        # the oracle's quantised symbols, recomputed here exactly as lteo_dci_decode does
        import ctypes as C
        tab = np.zeros(3 * K, np.uint16)
        S.lte_rm_conv_table.argtypes = [C.c_uint32, C.c_void_p]
        S.lte_rm_conv_table(K, ltelib.ptr(tab))
        rm = np.zeros(3 * K, np.float32)
        for k in range(E):
            rm[tab[k % (3 * K)]] = np.float32(rm[tab[k % (3 * K)]] + e[k])
        mx = np.float32(np.abs(rm).max())
        gain = np.float32(32.0) / mx
        v = np.clip(rm * gain + np.float32(127.5), 0.0, 255.0).astype(np.float32)
        r = (2 * v.astype(np.int64) - 255).reshape(3, K)
        mb, mc = model_decode(r, K, nb)
        assert np.array_equal(mb, bits) and mc == crc, (nb, trial, L)
