"""Host logic of the Coupled_FAST-compatible object (tropical_cyclone_risk_b200/coupled_fast.py) with a stand-in engine:
constructor attributes as the reference sets them (bam_track.py:51-60, coupled_fast.py:23-27), the basin crop of
init_fields, the phase draw order of gen_f, None for ventilated seeds, the OdeResult fields run_tracks reads."""
import datetime

import numpy as np


class FakeEngine:
    def __init__(self, p, device=0):
        self.p, self.n_steps, self.calls = p, int(p.n_steps), []

    def upload_static(self, st):
        self.static = st

    def alloc_tables(self, n, lon, lat):
        self.grid = (n, np.asarray(lon), np.asarray(lat))

    def upload_months(self, ym0, planes):
        self.planes = np.array(planes)

    def synchronize(self):
        pass

    def close(self):
        pass

    def integrate(self, ym, lon0, lat0, v0, m0, h_bl, phases):
        n = len(lon0)
        self.calls.append(dict(lon0=np.array(lon0), h_bl=np.array(h_bl), phases=np.array(phases)))
        track = np.full((n, self.n_steps, 4), np.nan)
        n_time = np.zeros(n, np.int32)
        status = np.zeros(n, np.int32)
        for i in range(n):
            if v0[i] < 0:                                                  # stand-in rule: negative v -> "ventilated"
                status[i] = 2
                continue
            n_time[i] = 5 + i
            track[i, :n_time[i]] = np.arange(n_time[i])[:, None] + np.array([lon0[i], lat0[i], v0[i], m0[i]])
            status[i] = 1
        return dict(track=track, n_time=n_time, status=status, nfev=np.full(n, 44, np.int32))


def test_shim_host_logic(monkeypatch):
    from tropical_cyclone_risk_b200 import compute, coupled_fast, engine, fields, layout, synth
    from tropical_cyclone_risk_b200 import namelist as nl
    monkeypatch.setattr(engine, "Engine", FakeEngine)
    import torch
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    lon, lat = synth.era5_axes()
    olon, olat = synth.ocean_axes()
    raw = synth.synth_month_raw(2001, 8, lon, lat)
    mld, strat = synth.synth_ocean(olon, olat, 8)
    _, _, pg = fields.prepare_month(nl, (0.0, -90.0, 360.0, 90.0), lon, lat, raw, olon, olat, mld, strat)
    wnd = dict(lon=lon, lat=lat, **{name: pg[i] for i, name in enumerate(layout.FIELD_NAMES[:14])})
    b = compute.TC_Basin("WP")
    fast = coupled_fast.Coupled_FAST(wnd, b, datetime.datetime(2001, 8, 15), 900, 15 * 86400, static=synth.synth_static(full_res=False))
    assert fast.total_steps == 1441 and fast.dt_track == 900 and fast.nWLvl == 4 and fast.nLvl == 2
    assert np.array_equal(fast.t_s, np.linspace(0, 15 * 86400, 1441)) and fast.h_bl == 1400.0
    assert abs(fast.beta - (1 - 0.33 - 0.1)) == 0.0
    fast.init_fields(lon, lat, pg[14], pg[15], pg[16], pg[17])
    eng = fast._engine
    assert eng.n_steps == 1441                                            # the constructor's interval, not the namelist's
    # the same planes the outer tier uploads for this basin-month (rh_mid excepted: the caller samples it itself)
    lon_b, lat_b, want = fields.prepare_month(nl, b.get_bounds(), lon, lat, raw, olon, olat, mld, strat)
    assert np.array_equal(eng.grid[1], lon_b) and np.array_equal(eng.grid[2], lat_b)
    assert np.array_equal(eng.planes[0, :18], want[:18]) and not eng.planes[0, 18].any()
    # gen_f draws rand(15, 1) once per series, series-major (bam_track.py:27, 111-113)
    monkeypatch.setattr(coupled_fast, "random_seed", lambda: np.random.seed(5))
    fast.h_bl = 1800.0
    res = fast.gen_track(140.0, 15.0, 5.0, 0.2)
    np.random.seed(5)
    want_ph = np.concatenate([np.random.rand(15, 1).reshape(-1) for _ in range(4)])
    assert np.array_equal(eng.calls[-1]["phases"][0], want_ph) and eng.calls[-1]["h_bl"][0] == 1800.0
    assert res.status == 1 and res.nfev == 44 and res.y.shape == (4, 5) and np.array_equal(res.t, fast.t_s[:5])
    assert np.array_equal(res.y[:, 0], [140.0, 15.0, 5.0, 0.2])
    assert fast.gen_track(140.0, 15.0, -1.0, 0.2) is None                 # coupled_fast.py:241-244
    out = fast.gen_tracks([140.0, 141.0], [15.0, 16.0], [5.0, -1.0], [0.2, 0.2])
    assert out[0] is not None and out[1] is None
