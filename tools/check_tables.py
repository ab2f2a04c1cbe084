"""Structural validation of include/lte_tables.h (run by tests/test_tables.py)."""
import re, sys, pathlib
import numpy as np

HDR = pathlib.Path(__file__).resolve().parents[1] / "include" / "lte_tables.h"

def _arr(txt, name):
    m = re.search(name + r"\[[^\]]*\]\s*(?:\[[^\]]*\])?\s*=\s*\{(.*?)\};", txt, re.S)
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    return [int(x) for x in re.findall(r"-?\d+", body)]

def load():
    txt = HDR.read_text()
    f1 = _arr(txt, "lte_qpp_f1"); f2 = _arr(txt, "lte_qpp_f2")
    tbs = _arr(txt, "lte_tbs_table")
    return f1, f2, tbs, _arr(txt, "lte_tbs_format1c")

def qpp_K(i):
    if i < 60: return 40 + 8 * i
    if i < 92: return 512 + 16 * (i - 59)
    if i < 124: return 1024 + 32 * (i - 91)
    return 2048 + 64 * (i - 123)

def segm(tbs):
    """36.212 5.1.2 -> (C, K+, K-, C+, C-, F)"""
    Ks = [qpp_K(i) for i in range(188)]
    B = tbs + 24
    if B <= 6144:
        C, Bp = 1, B
    else:
        C = -(-B // (6144 - 24)); Bp = B + C * 24
    Kp = min(k for k in Ks if C * k >= Bp)
    if C == 1:
        return C, Kp, 0, 1, 0, Kp - Bp
    Km = max(k for k in Ks if k < Kp)
    dK = Kp - Km
    Cm = (C * Kp - Bp) // dK
    Cp = C - Cm
    F = Cp * Kp + Cm * Km - Bp
    return C, Kp, Km, Cp, Cm, F

def check():
    f1, f2, tbs, f1c = load()
    errs = []
    assert len(f1) == 188 and len(f2) == 188, (len(f1), len(f2))
    for i in range(188):
        K = qpp_K(i)
        idx = np.arange(K, dtype=np.int64)
        p = (f1[i] * idx + f2[i] * idx * idx) % K
        if len(np.unique(p)) != K:
            errs.append(f"QPP row {i} K={K} f1={f1[i]} f2={f2[i]} not a bijection")
    if len(tbs) != 34 * 110:
        errs.append(f"TBS table has {len(tbs)} entries, expected {34*110}")
        return errs
    T = np.array(tbs).reshape(34, 110)
    for r in range(34):
        for c in range(110):
            v = int(T[r, c])
            if v % 8: errs.append(f"TBS[{r}][{c+1}]={v} not byte aligned")
            if segm(v)[5] != 0: errs.append(f"TBS[{r}][{c+1}]={v} has filler bits {segm(v)}")
            if c and T[r, c] < T[r, c - 1]: errs.append(f"TBS[{r}][{c+1}]={v} < left {T[r,c-1]}")
            if r and r < 27 and T[r, c] < T[r - 1, c]: errs.append(f"TBS[{r}][{c+1}]={v} < above {T[r-1,c]}")
    return errs

if __name__ == "__main__":
    e = check()
    print("\n".join(e) if e else "tables OK")
    sys.exit(1 if e else 0)
