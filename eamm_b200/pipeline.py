"""Host<->device pipelining for clip generation.

The reference's caller generates a clip frame by frame and copies every frame back to the host
(`demo.py:251-281`: `out['prediction'].data.cpu().numpy()` per frame).  `FramePipeline` keeps that
contract (host tensors in, host tensors out, every batch copied both ways) but overlaps the copies
with compute on three CUDA streams: H2D of batch i+1 | generator forward of batch i | D2H of batch i-1.
Nothing is skipped or cached: each submitted batch is uploaded, generated and downloaded.
"""
import torch


class FramePipeline:
    def __init__(self, generator, depth=2):
        p = next(generator.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("eamm_b200: FramePipeline needs the generator on a CUDA device")
        self.gen, self.dev, self.depth = generator, p.device, depth
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_run = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self.in_ready = [torch.cuda.Event() for _ in range(depth)]
        self.run_done = [torch.cuda.Event() for _ in range(depth)]
        self.staging = [None] * depth
        self.count = 0
        self._strict = generator.strict_errors
        generator.strict_errors = False          # the singular-Jacobian flag is checked in drain()

    def _stage(self, slot, h_src, h_kpd, h_kps):
        st = self.staging[slot]
        if st is None or st[0].shape != h_src.shape:
            mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device=self.dev)
            st = (mk(h_src), {k: mk(v) for k, v in h_kpd.items()}, {k: mk(v) for k, v in h_kps.items()})
            self.staging[slot] = st
        st[0].copy_(h_src, non_blocking=True)
        for k, v in h_kpd.items():
            st[1][k].copy_(v, non_blocking=True)
        for k, v in h_kps.items():
            st[2][k].copy_(v, non_blocking=True)
        return st

    def submit(self, h_src, h_kpd, h_kps, h_out, key="prediction"):
        """Queue one batch: pinned host inputs -> `h_out` (pinned host tensor shaped like out[key])."""
        slot = self.count % self.depth
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(self.run_done[slot])       # the forward that last read this slot is done
            d_src, d_kpd, d_kps = self._stage(slot, h_src, h_kpd, h_kps)
            self.in_ready[slot].record(self.s_in)
        with torch.cuda.stream(self.s_run):
            self.s_run.wait_event(self.in_ready[slot])
            out = self.gen(d_src, kp_driving=d_kpd, kp_source=d_kps)
            res = out[key]
            self.run_done[slot].record(self.s_run)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.run_done[slot])
            h_out.copy_(res, non_blocking=True)
            res.record_stream(self.s_out)
        self.count += 1

    def drain(self):
        """Block until every submitted batch has landed in its host buffer; re-raise device-side errors."""
        self.s_out.synchronize()
        self.s_run.synchronize()
        eng = getattr(self.gen, "_eng", None)
        if self._strict and eng is not None and getattr(eng, "dm", None) is not None:
            # the flag is sticky across the non-strict forwards queued since the last drain: a singular driving
            # Jacobian in ANY of those batches is reported here (the reference raises from torch.inverse)
            from .modules.dense_motion import check_status
            check_status(eng.dm.status, clear=True)

    def close(self):
        self.drain()
        self.gen.strict_errors = self._strict
