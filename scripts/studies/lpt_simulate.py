"""Lane-level list-scheduling model of the persistent integrator (56 832 lanes, one storm per lane, storms popped
from a global queue, duration = RK attempts): makespan for different queue orders.  Input: lpt_collect.py."""
import numpy as np, heapq, time
d = np.load("/tmp/lpt_data.npz"); F, D = d["F"], d["D"]
steps = np.ceil(np.maximum(D, 1) / 6.0).astype(np.int64)       # macro steps per storm (one RK attempt = 6 RHS)
M = 148 * 2 * 6 * 32
def makespan(order):
    # lanes pop storms in `order`; event-driven: heap of (free_time, lane)
    h = [0] * M
    heapq.heapify(h)
    end = 0
    for s in steps[order]:
        t = heapq.heappop(h)
        t2 = t + int(s)
        if t2 > end: end = t2
        heapq.heappush(h, t2)
    return end
n = D.size
ideal = steps.sum() / M
rng = np.random.default_rng(0)
print("storms", n, "lanes", M, "ideal (perfect balance)", round(ideal, 1), "longest storm", steps.max())
print("attempt order   ", makespan(np.arange(n)))
print("random order    ", makespan(rng.permutation(n)))
print("true LPT        ", makespan(np.argsort(-steps, kind="stable")))
# predictors from seeding-time features: lat, v0, m0, vpot, chi, rh, shear, bathy, lon
names = ["lat", "v0", "m0", "vpot", "chi", "rh", "shear", "bathy", "lon"]
for k, nm in enumerate(names):
    c = np.corrcoef(F[:, k], D)[0, 1]
    print("corr(nfev, %s) = %.3f" % (nm, c))
# linear + quadratic regression on half, evaluate ordering on all
X = np.column_stack([F, F**2, np.ones(n)])
X = (X - X.mean(0)) / (X.std(0) + 1e-12); X[:, -1] = 1
half = rng.random(n) < 0.5
w, *_ = np.linalg.lstsq(X[half], D[half], rcond=None)
pred = X @ w
print("R^2 of the quadratic fit (held-out): %.3f" % (1 - ((D[~half] - pred[~half])**2).sum() / ((D[~half] - D[~half].mean())**2).sum()))
print("predicted-LPT   ", makespan(np.argsort(-pred, kind="stable")))
# coarse: two classes only (top 20 % predicted first)
thr = np.quantile(pred, 0.8)
order2 = np.concatenate([np.flatnonzero(pred >= thr), np.flatnonzero(pred < thr)])
print("top-20%-first   ", makespan(order2))
