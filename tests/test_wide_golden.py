"""The CPU oracle (oracle/tcr_oracle.c) against the wide fixtures of the UNMODIFIED reference: 510 storms over
five basin-months -- southern-hemisphere genesis on the global grid, the 0/360 E seam, all five boundary-layer
depths, 900-second output (1441 samples).  Rules and report: tests/wide.py."""
import numpy as np
import pytest

import wide
from oracle import tcr_oracle as orc


@pytest.mark.parametrize("name", wide.WIDE_CASES)
def test_oracle_vs_wide_reference(name):
    g, case = wide.load_case(name)
    o = orc.integrate_batch(case.p, case.env, *wide.seeds_of(g), post_all=True, n_threads=8)
    rep = wide.check(name, g, o, o["n_clean"])
    # the reference-stable majority carries the pin; the sensitive storms are the reference's own, not ours
    assert rep["reference_stable"] >= 0.6 * rep["storms"]
    assert rep["worst_stable_err"] < 1e-5


def test_wide_fixtures_cover_what_they_claim():
    tot = 0
    hbl = set()
    for name in wide.WIDE_CASES:
        g = wide.golden("ref_wide_%s.npz" % name)
        tot += g["lon0"].size
        hbl |= set(np.unique(g["h_bl"]).tolist())
    assert tot >= 500
    assert hbl == {1400.0, 1500.0, 1600.0, 1800.0, 2000.0}
    g = wide.golden("ref_wide_gl_feb.npz")
    assert (g["lat0"] < 0).all() and str(g["basin"]) == "GL"
    g = wide.golden("ref_wide_wp_aug_900.npz")
    assert int(g["interval"]) == 900 and int(g["n_time"].max()) > 361
    g = wide.golden("ref_wide_gl_sep.npz")
    assert (g["lon0"] > 355).any() and (g["lat0"] < 0).any() and (g["lat0"] > 40).any()
