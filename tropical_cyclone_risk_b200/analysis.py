"""Return-period / landfall analysis on the finished-track tensor ("next" row N4):
what notebooks/sample_analysis.ipynb cells 13-17 do with xarray, on the GPU and across ranks.

    vmax_at_poi = where(haversine(poi, track) <= radius, vmax_trks).max(time)      (cell 15)
    exceedance_count[b] = sum(vmax_at_poi >= vmax_bins[b])                        (cell 17)
    return_period = total_years / exceedance_count                                (cell 17)

Tracks are sharded by rows over torch.distributed ranks (no data-path collective); the only
exchange is the all-reduce of the per-bin counts.
"""
import numpy as np

R_EARTH_NOTEBOOK_M = 6378000.0            # cell 13 (the model itself uses 6378100 m, util/constants.py:7)
DEFAULT_BINS = np.arange(10, 81, 5)       # cell 17


def vmax_at_poi(engine, lon_trks, lat_trks, vmax_trks, poi_lon, poi_lat, radius_km=100.0):
    """Per-track maximum intensity while within radius_km of the point of interest (NaN if never)."""
    return engine.poi_vmax(lon_trks, lat_trks, vmax_trks, poi_lon, poi_lat, radius_km, R_EARTH_NOTEBOOK_M)


def exceedance_counts(engine, vmax_poi, bins=DEFAULT_BINS, reduce_over_ranks=True):
    """Number of tracks reaching each bin, summed over all ranks when torch.distributed is initialised."""
    counts = engine.exceedance(vmax_poi, bins)
    if reduce_over_ranks:
        counts = all_reduce_counts(counts)
    return counts


def all_reduce_counts(counts):
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return counts
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts
    t = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.int64))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def return_period(counts, total_years):
    """Years between exceedances of each bin (inf where never exceeded)."""
    counts = np.asarray(counts, dtype=np.float64)
    with np.errstate(divide="ignore"):
        return np.where(counts > 0, float(total_years) / counts, np.inf)
