"""The product's RNTI history (ltesniffer_b200/csrc/host_search.cpp) against the REFERENCE'S OWN RNTIManager,
compiled from /root/reference/lib/src/util/*.cc into oracle/_ref (random operation sequences)."""
import numpy as np
import pytest
import ltelib
from ltesniffer_b200 import capi

pytestmark = pytest.mark.skipif(not (ltelib.ref_available() or __import__("os").path.isdir("/root/reference")), reason="oracle/_ref not built")


@pytest.mark.parametrize("seed,nrnti,steps", [(1, 40, 1500), (2, 400, 700), (3, 6, 12500)])
def test_matches_reference_rnti_manager(infra, seed, nrnti, steps):
    R = ltelib.rntimgr_ref()
    ref = R.rnti_manager_create(9, 304 // 5, 5)
    s = capi.Search(100, 2, 1, 1, threshold=5)   # seeds the same ranges as LTESniffer_Core.cc:398-417
    for f in (2, 4):
        R.rnti_manager_add_evergreen(ref, 1, 10, f)
        R.rnti_manager_add_evergreen(ref, 0xFFFE, 0xFFFF, f)
    for f in range(9):
        R.rnti_manager_add_forbidden(ref, 0, 0, f)
    L = s.L
    rng = np.random.default_rng(seed)
    pool = np.concatenate([rng.integers(11, 0xFFF3, nrnti), [0, 1, 5, 10, 0xFFFE, 0xFFFF, 0xFFF5]]).astype(np.uint16)
    for step in range(steps):
        for _ in range(int(rng.integers(0, 25))):
            r = int(rng.choice(pool))
            f = int(rng.integers(0, 9))
            op = rng.random()
            if op < 0.55:
                R.rnti_manager_add_candidate(ref, r, f)
                L.ltephy_search_rnti_add_candidate(s.h, r, f)
            elif op < 0.95:
                a, b = R.rnti_manager_validate_and_refresh(ref, r, f), L.ltephy_search_rnti_validate_and_refresh(s.h, r, f)
                assert bool(a) == bool(b), (step, r, f)
            else:
                R.rnti_manager_activate_and_refresh(ref, r, f, 3)
                s.L.ltephy_search_activate(s.h, r, f, 3)
            assert R.rnti_manager_getFrequency(ref, r, f) == L.ltephy_search_rnti_frequency(s.h, r, f)
            assert R.rnti_manager_get_associated_format_idx(ref, r) == L.ltephy_search_rnti_assoc_format(s.h, r)
            assert R.rnti_manager_get_activation_reason(ref, r) == L.ltephy_search_rnti_reason(s.h, r)
            assert bool(R.rnti_manager_is_forbidden(ref, r, f)) == bool(L.ltephy_search_rnti_is_forbidden(s.h, r, f))
            assert bool(R.rnti_manager_is_evergreen(ref, r, f)) == bool(L.ltephy_search_rnti_is_evergreen(s.h, r, f))
        R.rnti_manager_step_time(ref)
        L.ltephy_search_rnti_step_time(s.h)
    R.rnti_manager_free(ref)
