"""CPU study (oracle only, no GPU): RHS count of every integrated storm of a configs[1] step together with what
is known about it at seeding time -> /tmp/lpt_data.npz (input of lpt_simulate.py).  Run from the repo root."""
import sys, time, heapq, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from conftest import Case
from oracle import tcr_oracle as orc
years = list(range(2001, 2011))
case = Case("NA", years)
feat = []; dur = []
t0 = time.time()
for yi, y in enumerate(years):
    r = orc.run_attempts(case.p, case.env, 12*yi, case.masks, 20260101, y, 0, 362000, want_tracks=False, n_threads=8)
    sel = r["nfev"] > 0          # integrated (status != vent with nfev 0 also integrated but trivial)
    integ = (r["code"] == 2) if (r["code"] == 2).any() else sel
    ic = r["ic"][integ]; month = r["month"][integ]
    q = orc.env_interp(case.env, (12*yi + month - 1).astype(np.int32), ic[:,0], ic[:,1])
    f = np.column_stack([ic[:,1], ic[:,2], ic[:,3], q[:,15], q[:,14], q[:,18], np.hypot(q[:,0]-q[:,2], q[:,1]-q[:,3]), q[:,19], ic[:,0]])
    feat.append(f); dur.append(r["nfev"][integ])
    print(y, integ.sum(), "integrated, mean nfev %.1f" % r["nfev"][integ].mean(), "%.0fs" % (time.time()-t0), flush=True)
F = np.vstack(feat); D = np.concatenate(dur).astype(np.float64)
np.savez("/tmp/lpt_data.npz", F=F, D=D)
print("storms", D.size, "mean", D.mean(), "p50", np.median(D), "p90", np.percentile(D,90), "max", D.max(), "zero", (D==0).mean())
