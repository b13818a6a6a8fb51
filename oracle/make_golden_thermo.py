"""Golden vectors of the thermodynamic pre-processing from the UNMODIFIED reference:
thermo.CAPE_PI_vectorized / sat_deficit / conv_q_to_rh exactly as thermo/calc_thermo.py:60-69 calls
them, on synthetic ERA5-shaped soundings.  Build container only (needs /root/reference):

    python oracle/make_golden_thermo.py                ->  tests/golden/ref_thermo.npz
    python oracle/make_golden_thermo.py --reversible   ->  tests/golden/ref_thermo_rev.npz   (namelist.select_thermo = 2)
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh                                  # noqa: E402
from tropical_cyclone_risk_b200 import synth_thermo                   # noqa: E402

N = 2048
K_MID = 13                                                            # 600 hPa = namelist.p_midlevel (nearest level)


def main():
    if not rh.available():
        raise SystemExit("reference tree not present")
    ref = rh.load_reference()
    th, nl = ref.thermo, ref.namelist
    assert nl.select_thermo == 1 and nl.select_interp == 2
    p, ta, hus, sst, psl = synth_thermo.soundings(N, seed=7)
    assert p[K_MID] == nl.p_midlevel
    nlat, nlon = 32, N // 32
    shp = (nlat, nlon)
    ta64 = ta.astype(np.float64).reshape((p.size,) + shp)
    hus64 = hus.astype(np.float64).reshape((p.size,) + shp)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        vmax = th.CAPE_PI_vectorized(sst.reshape(shp), psl.reshape(shp), p.copy(), ta64, hus64)       # calc_thermo.py:60-61
        chi = np.minimum(np.maximum(th.sat_deficit(sst.reshape(shp), psl.reshape(shp), ta64[K_MID], float(p[K_MID]),
                                                   hus64[K_MID]), 0), 10)                                # :66-68
        rh_mid = th.conv_q_to_rh(ta64[K_MID], hus64[K_MID], float(p[K_MID]))                            # :69
    # the entropy inversion table CAPE_PI_vectorized loaded (thermo.py:274-278) travels inside the fixture: it is an
    # input of the computation, and the GPU box has no reference tree to read it from
    with np.load(os.path.join(rh.REF_ROOT, "thermo", "entropy_table.npz")) as t:
        tp, ts, tT = np.array(t["p"]), np.array(t["s"]), np.array(t["T"])
        crc = int(np.frombuffer(tT.tobytes(), dtype=np.uint32).sum() & 0xffffffff)
    out = os.path.join(ROOT, "tests", "golden", "ref_thermo.npz")
    np.savez_compressed(out, p=p, ta=ta, hus=hus, sst=sst, psl=psl, k_mid=K_MID, cecd=nl.Ck / nl.Cd,
                        vmax=vmax.reshape(-1), chi=chi.reshape(-1), rh_mid=rh_mid.reshape(-1), table_crc=crc,
                        table_p=tp, table_s=ts, table_T=tT)
    print("wrote", out, os.path.getsize(out), "bytes; PI>0 in %d of %d columns, max %.1f m/s" %
          ((vmax > 0).sum(), N, np.nanmax(vmax)))


def main_reversible():
    """namelist.select_thermo = 2 (reversible thermodynamics; thermo.py:56-60, 71-75, 132-133, 279-284, 343-353): the same
    soundings through the unmodified reference with the switch flipped.  The reference loads its 8 MB inversion table
    thermo/entropy_table_reversible.npz from namelist.src_directory; the fixture carries every third pressure / entropy node
    and every fourth total-water node of it (34 x 34 x 25), written to a scratch directory the namelist is pointed at, so
    that the table the reference interpolated in is exactly the one inside the fixture.  -> tests/golden/ref_thermo_rev.npz"""
    import tempfile
    ref = rh.load_reference()
    th, nl = ref.thermo, ref.namelist
    with np.load(os.path.join(rh.REF_ROOT, "thermo", "entropy_table_reversible.npz")) as t:
        tp, ts, tr, tT = np.array(t["p"][::3]), np.array(t["s"][::3]), np.array(t["rt"][::4]), np.array(t["T"][::3, ::3, ::4])
    p, ta, hus, sst, psl = synth_thermo.soundings(N, seed=11)
    nlat, nlon = 32, N // 32
    shp = (nlat, nlon)
    ta64 = ta.astype(np.float64).reshape((p.size,) + shp)
    hus64 = hus.astype(np.float64).reshape((p.size,) + shp)
    saved = (nl.select_thermo, nl.src_directory)
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "thermo"))
        np.savez(os.path.join(tmp, "thermo", "entropy_table_reversible.npz"), p=tp, s=ts, rt=tr, T=tT)
        nl.select_thermo, nl.src_directory = 2, tmp
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                vmax = th.CAPE_PI_vectorized(sst.reshape(shp), psl.reshape(shp), p.copy(), ta64, hus64)
                chi = np.minimum(np.maximum(th.sat_deficit(sst.reshape(shp), psl.reshape(shp), ta64[K_MID], float(p[K_MID]),
                                                           hus64[K_MID]), 0), 10)
                rh_mid = th.conv_q_to_rh(ta64[K_MID], hus64[K_MID], float(p[K_MID]))
        finally:
            nl.select_thermo, nl.src_directory = saved
    out = os.path.join(ROOT, "tests", "golden", "ref_thermo_rev.npz")
    np.savez_compressed(out, p=p, ta=ta, hus=hus, sst=sst, psl=psl, k_mid=K_MID, cecd=nl.Ck / nl.Cd,
                        vmax=vmax.reshape(-1), chi=chi.reshape(-1), rh_mid=rh_mid.reshape(-1),
                        table_p=tp, table_s=ts, table_rt=tr, table_T=tT)
    print("wrote", out, os.path.getsize(out), "bytes; PI>0 in %d of %d columns, max %.1f m/s" %
          ((vmax > 0).sum(), N, np.nanmax(vmax)))


if __name__ == "__main__":
    if "--reversible" in sys.argv:
        main_reversible()
    else:
        main()
