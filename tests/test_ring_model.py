"""The residency rule of the integrator's Fourier rings, checked on a host model of the protocol.

k_integrate (csrc/tcr_kernels.cuh: ring_need / ring_serve / the CTA request list, tcr_ring_fill) keeps, per storm, only the
last two SEGMENTS of its tabulated Fourier nodes: node j lives in ring slot j & (ring_nodes - 1), `have` says that nodes
[0, have) have been tabulated so far, and a segment is tabulated when the next RK attempt could bracket a node >= have.
The bits the kernel computes are compared with the full tables on the GPU (tests/test_gpu_parity.py::test_fourier_ring_*);
what this file checks, without a GPU, is the ARGUMENT that no evaluation ever reads a node that has been overwritten or
not been written yet -- for the segment length the host chooses (tcrisk.cu: ring_nodes_for), for any sequence of accepted
and rejected steps up to max_step, for an initial-step probe (select_initial_step's second evaluation at t = h0) that lands
anywhere in the record, and for stage times that round one ulp past the end of the attempt.
"""
import math

import numpy as np
import pytest

RK_C = (0.2, 0.3, 0.8, 8.0 / 9.0, 1.0, 1.0)          # stage abscissae of RK45 + the FSAL evaluation at t + h


def ring_nodes_for(n_steps, t_step, max_step, forced=True):
    """tcrisk.cu: ring_nodes_for (TCR_FTAB_RING=1: the >= 4 rings rule of the default is not applied)."""
    span = math.ceil(max_step / t_step) + 1.0
    seg = 64
    while seg < span + 4.0 and seg < (1 << 20):
        seg *= 2
    if 2 * seg >= n_steps:
        return 0
    if not forced and 8 * seg > n_steps:
        return 0
    return 2 * seg


class Ring:
    """One storm's ring: which node each slot holds."""

    def __init__(self, ring_nodes, n_steps):
        self.n, self.ns = ring_nodes, n_steps
        self.slot = np.full(ring_nodes, -1, dtype=np.int64)
        self.have = 0
        self.fills = 0

    def fill(self):                                   # tcr_ring_fill: nodes [have, have + seg)
        seg = self.n // 2
        for j in range(self.have, self.have + seg):
            if j < self.ns:
                self.slot[j & (self.n - 1)] = j
        self.have += seg
        self.fills += 1

    def serve(self, need):                            # ring_serve (and the one-segment post of the CTA list followed by it)
        while need >= self.have:
            self.fill()

    def read(self, j):
        assert self.slot[j & (self.n - 1)] == j, "node %d read, slot holds %d (have %d)" % (j, self.slot[j & (self.n - 1)], self.have)


def fs_index(t_s, t):
    """tcr_fs_index: searchsorted(t_s, t, 'left') clipped to [1, n - 1]."""
    return int(min(max(np.searchsorted(t_s, t, side="left"), 1), len(t_s) - 1))


def ring_need(idx, ns):
    return idx if idx < 0 else min(idx + 1, ns - 1)


def run_storm(rng, t_s, max_step, ring_nodes, h0_far):
    ns, T = len(t_s), t_s[-1]
    ring = Ring(ring_nodes, ns)

    def evaluate(te):
        idx = fs_index(t_s, te)
        ring.read(idx - 1)
        ring.read(idx)

    # M_INIT0: the evaluation at t = 0 brackets nodes 0 and 1
    ring.serve(ring_need(1, ns))
    evaluate(0.0)
    # M_INIT1: select_initial_step's probe at t = h0 -- seconds to minutes in practice, anywhere in the record here
    h0 = rng.uniform(0.0, T) if h0_far else rng.uniform(1e-6, 600.0)
    ring.serve(ring_need(fs_index(t_s, 0.0 + h0), ns))
    evaluate(0.0 + h0)
    if ring.have > ring.n:                            # the probe overwrote the nodes the integration starts from
        ring.have = 0
    t, h_abs = 0.0, min(max_step, rng.uniform(1.0, max_step))
    n_eval = 2
    while t < T and n_eval < 4000:
        h_abs = min(h_abs, max_step)
        t_new = t + h_abs
        if t_new - T > 0.0:
            t_new = T
        h = t_new - t
        ring.serve(ring_need(fs_index(t_s, t_new), ns))
        for c in RK_C:
            evaluate(t + c * h)
            n_eval += 1
        # a stage time one ulp beyond the end of the attempt (t + 1.0 * h need not equal t_new)
        evaluate(np.nextafter(t_new, np.inf) if t_new < T else T)
        if rng.random() < 0.15:                       # rejected: the step shrinks, the clock stays
            h_abs *= rng.uniform(0.2, 0.9)
        else:
            t = t_new
            h_abs *= rng.uniform(0.3, 10.0)
            if rng.random() < 0.01:                   # the storm ends
                break
    return ring.fills


@pytest.mark.parametrize("interval,max_step", [(3600, 86400.0), (900, 86400.0), (3600, 6 * 3600.0), (1800, 86400.0), (600, 43200.0)])
def test_every_evaluation_reads_resident_nodes(interval, max_step):
    T = 15 * 86400.0
    n_steps = int(T / interval) + 1
    t_s = np.linspace(0.0, T, n_steps)
    ring_nodes = ring_nodes_for(n_steps, T / (n_steps - 1), max_step)
    assert ring_nodes > 0 and ring_nodes & (ring_nodes - 1) == 0
    rng = np.random.default_rng(interval)
    fills = [run_storm(rng, t_s, max_step, ring_nodes, h0_far=(i % 7 == 0)) for i in range(300)]
    # storms that cross the whole record tabulate every segment once (plus the re-tabulation after a far probe)
    assert max(fills) >= math.ceil(n_steps / (ring_nodes // 2))


def test_a_shorter_segment_would_fail():
    """The rule is tight enough to matter: with segments no longer than what one attempt spans, a read misses."""
    T, interval, max_step = 15 * 86400.0, 900, 86400.0
    n_steps = int(T / interval) + 1
    t_s = np.linspace(0.0, T, n_steps)
    rng = np.random.default_rng(1)
    with pytest.raises(AssertionError):
        for _ in range(200):
            run_storm(rng, t_s, max_step, 128, h0_far=False)          # segments of 64 nodes, attempts of up to 96


def test_ring_default_rule():
    """Rings are the default where the grid is at least four rings long (900-s output), not at 361 hourly nodes."""
    T = 15 * 86400.0
    assert ring_nodes_for(361, T / 360, 86400.0, forced=False) == 0
    assert ring_nodes_for(361, T / 360, 86400.0, forced=True) == 128
    assert ring_nodes_for(1441, T / 1440, 86400.0, forced=False) == 256
    assert ring_nodes_for(100, T / 99, 86400.0, forced=True) == 0       # a ring as long as the table saves nothing
