"""Double-buffered year batches: while batch i's finished tracks travel to the host (copy engine,
side stream), batch i+1 is already being integrated (SMs, main stream).

The reference's `run_downscaling` (util/compute.py:216-242) collects one 9-tuple per year from its
dask workers; here a *batch* of years is one `tcr_run_years` call that leaves its result block in
HBM, and the device->host copy of that block (72 B per track sample) is the only part of the
end-to-end path that does not need the SMs -- so it is overlapped with the next batch.

torch is used for what it is good at here: device / pinned host memory and streams.
"""
import numpy as np

from . import layout


class _Block:
    """One result block: flat float64 buffer on the device, pinned twin on the host, and the
    section offsets of the 9-tuple inside it."""

    def __init__(self, torch, dev, n_years, n_tracks, n_steps, pinned=True):
        rows = n_years * n_tracks
        self.shapes = [("lon", (n_years, n_tracks, n_steps)), ("lat", (n_years, n_tracks, n_steps)),
                       ("v", (n_years, n_tracks, n_steps)), ("m", (n_years, n_tracks, n_steps)),
                       ("vmax", (n_years, n_tracks, n_steps)), ("env", (n_years, n_tracks, n_steps, 4)),
                       ("tc_month", (n_years, n_tracks)), ("n_seeds", (n_years, len(layout.BASIN_IDS), 12))]
        sizes = [int(np.prod(s)) for _, s in self.shapes] + [(rows + 1) // 2]        # tc_basin: int32 pairs
        self.offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        total = int(self.offsets[-1])
        self.dev = torch.empty(total, dtype=torch.float64, device=dev)
        self.host = torch.empty(total, dtype=torch.float64).pin_memory() if pinned else None
        self.rows = rows
        self.dptr = {name: self.dev.data_ptr() + int(self.offsets[i]) * 8 for i, (name, _) in enumerate(self.shapes)}
        self.dptr["tc_basin"] = self.dev.data_ptr() + int(self.offsets[len(self.shapes)]) * 8
        self.copied = torch.cuda.Event()
        self.busy = False
        self.nbytes = total * 8

    def host_views(self):
        return self.views_of(self.host.numpy())

    def views_of(self, h):
        """The sections of a flat float64 host array laid out like this block."""
        out = {name: h[int(self.offsets[i]):int(self.offsets[i + 1])].reshape(shape)
               for i, (name, shape) in enumerate(self.shapes)}
        b0 = int(self.offsets[len(self.shapes)])
        out["tc_basin"] = h[b0:].view(np.int32)[:self.rows].reshape(self.shapes[6][1])
        return out

    def copy_prefix_from(self, sub):
        """Device copy of a block with fewer year slots into the leading year slots of every section of this one."""
        for i, (name, shape) in enumerate(self.shapes):
            n = int(np.prod(sub.shapes[i][1]))
            self.dev[int(self.offsets[i]):int(self.offsets[i]) + n].copy_(sub.dev[int(sub.offsets[i]):int(sub.offsets[i]) + n])
        b0, s0 = int(self.offsets[len(self.shapes)]), int(sub.offsets[len(sub.shapes)])
        n = (sub.rows + 1) // 2
        self.dev[b0:b0 + n].copy_(sub.dev[s0:s0 + n])


class YearPipeline:
    """submit() runs a batch of years and starts its download; result() hands out the host arrays.

    depth result blocks are cycled; submit() blocks only if the block it needs is still being
    downloaded from `depth` submissions ago."""

    def __init__(self, engine, n_years, n_tracks, depth=2, device=None):
        import torch
        self.torch = torch
        self.eng = engine
        self.dev = torch.device("cuda", engine.device if device is None else device)
        self.main = torch.cuda.current_stream(self.dev)
        self.copy = torch.cuda.Stream(device=self.dev)
        self.blocks = [_Block(torch, self.dev, n_years, n_tracks, engine.n_steps) for _ in range(depth)]
        self.n_tracks = n_tracks
        self.count = 0
        engine.set_stream(self.main.cuda_stream)

    @property
    def d2h_bytes(self):
        return self.blocks[0].nbytes

    def submit(self, ym_base, year_key, run_seed):
        blk = self.blocks[self.count % len(self.blocks)]
        if blk.busy:
            blk.copied.synchronize()                     # its previous download must have left the device block
        stats = self.eng.run_years_dev(ym_base, year_key, run_seed, self.n_tracks, blk.dptr)    # synchronous on main
        self.copy.wait_stream(self.main)
        with self.torch.cuda.stream(self.copy):
            blk.host.copy_(blk.dev, non_blocking=True)
            blk.copied.record(self.copy)
        blk.busy = True
        ticket = self.count
        self.count += 1
        return ticket, stats

    def upload_tables_async(self, ym0, planes):
        """Start uploading the NEXT batch's monthly planes (pinned host memory, [n][19][nlat][nlon]) into table
        slots [ym0, ym0 + n) on the copy stream, while the batch submitted next still computes on the slots it
        was given.  The batch that will use these slots must be submitted after `tables_ready()`."""
        self.copy.wait_stream(self.main)                 # the slots' previous readers have been enqueued (and finished)
        self.eng.set_stream(self.copy.cuda_stream)
        try:
            self.eng.upload_months(ym0, planes)
        finally:
            self.eng.set_stream(self.main.cuda_stream)
        self._tables_event = self.torch.cuda.Event()
        self._tables_event.record(self.copy)

    def tables_ready(self):
        """Order the main stream after the last asynchronous table upload."""
        ev = getattr(self, "_tables_event", None)
        if ev is not None:
            self.main.wait_event(ev)
            self._tables_event = None

    def result(self, ticket):
        """Host arrays (views of pinned memory, valid until `depth` further submissions)."""
        blk = self.blocks[ticket % len(self.blocks)]
        blk.copied.synchronize()
        return blk.host_views()

    def drain(self):
        for blk in self.blocks:
            if blk.busy:
                blk.copied.synchronize()
