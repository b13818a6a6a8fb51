import os
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
GLOBAL_BOUNDS = (0.0, -90.0, 360.0, 90.0)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a usable CUDA device: skip them (with the reason) instead of erroring in tcr_create."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    try:
        import torch
        ok = torch.cuda.is_available()
    except Exception:
        ok = False
    if not ok:
        skip = pytest.mark.skip(reason="no usable CUDA device (the hot path has no CPU fallback)")
        for it in gpu_items:
            it.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def _workload_cls():
    from tropical_cyclone_risk_b200.workload import Workload
    return Workload


class Case(_workload_cls()):
    """One basin's prepared inputs (tropical_cyclone_risk_b200.workload.Workload) plus the same
    numbers in the CPU oracle's layout."""

    def __init__(self, basin, years, months=range(1, 13), full_res=False, roughness=1.0,
                 zero_cov_over_land=False, namelist=None):
        from oracle import tcr_oracle as orc
        super().__init__(basin, years, months, full_res, roughness, zero_cov_over_land, namelist)
        self.env = orc.OracleEnv(self.lon, self.lat, self.planes, self.static)
        self.masks = orc.Masks(self.mask_lon, self.mask_lat, self.mask_planes)


@pytest.fixture(scope="session")
def na_case():
    """NA, year 2000, September only -- the inputs the golden fixtures were generated on."""
    from tropical_cyclone_risk_b200 import fields, synth
    from tropical_cyclone_risk_b200 import namelist as nl
    c = Case("NA", [2000], months=[9])
    lon, lat = synth.era5_axes()
    olon, olat = synth.ocean_axes()
    raw = synth.synth_month_raw(2000, 9, lon, lat)
    mld, strat = synth.synth_ocean(olon, olat, 9)
    _, _, pg = fields.prepare_month(nl, GLOBAL_BOUNDS, lon, lat, raw, olon, olat, mld, strat)
    c.planes_crc = zlib.crc32(np.ascontiguousarray(pg).tobytes())
    return c


@pytest.fixture(scope="session")
def na_year():
    """NA, year 2001, twelve months (seeding / run_year tests)."""
    return Case("NA", [2001])


@pytest.fixture(scope="session")
def gl_year():
    return Case("GL", [2002])
