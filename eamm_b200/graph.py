"""CUDA-graph replay of the generator for the reference's real access pattern: one frame (or a small,
fixed batch) per call with launch overhead larger than the work (`demo.py:251-281` calls the
generator with batch 1 for every frame).  The graph holds exactly the kernels an eager call launches;
inputs are copied into static buffers, outputs are static tensors overwritten by the next replay."""
import torch


class GraphedGenerator:
    def __init__(self, generator, source_image, kp_driving, kp_source, warmup=3, fixed_source=False):
        """Capture `generator(source_image, kp_driving, kp_source)` for these shapes (CUDA tensors).

        fixed_source=True is the clip case (one source image, many driving frames): the encoder and the
        anti-aliased copy are computed once here, the graph holds only the per-frame kernels, and
        `__call__` ignores its `source_image` argument."""
        p = next(generator.parameters())
        if p.device.type != "cuda":
            raise RuntimeError("eamm_b200: GraphedGenerator needs the generator on a CUDA device")
        self.gen = generator
        self.s_src = source_image.clone()
        self.s_kpd = {k: v.clone() for k, v in kp_driving.items() if torch.is_tensor(v)}
        self.s_kps = {k: v.clone() for k, v in kp_source.items() if torch.is_tensor(v)}
        strict = generator.strict_errors
        cache = getattr(generator, "cache_source", False)
        self.fixed_source = fixed_source
        generator.strict_errors = False                  # no host read inside a capture
        # fixed source: warm-up fills the encoder cache and the capture hits it (the graph holds no encoder kernels);
        # otherwise the cache must be OFF while capturing, or a caller that had enabled it would get a graph without
        # the encoder and stale features after `s_src.copy_(new_source)`
        generator.cache_source = bool(fixed_source)
        try:
            side = torch.cuda.Stream(p.device)
            side.wait_stream(torch.cuda.current_stream(p.device))
            with torch.cuda.stream(side):
                for _ in range(warmup):                  # plans, weight packings, workspaces, func attributes
                    generator(self.s_src, kp_driving=self.s_kpd, kp_source=self.s_kps)
            torch.cuda.current_stream(p.device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = generator(self.s_src, kp_driving=self.s_kpd, kp_source=self.s_kps)
        finally:
            generator.strict_errors = strict
            generator.cache_source = cache
        self._strict = strict
        from . import engine
        self._workspaces = list(engine._SPLITK_WS.values())     # the captured kernels point into these
        eng = generator._eng                                    # ... and into the per-shape workspaces (bounded caches)
        self._ws_refs = list(eng.ws.values()) + (list(eng.dm.ws.values()) if eng.dm is not None else [])

    def __call__(self, source_image, kp_driving, kp_source, check=True):
        if not self.fixed_source and source_image is not self.s_src:
            self.s_src.copy_(source_image, non_blocking=True)
        for k, v in self.s_kpd.items():
            v.copy_(kp_driving[k], non_blocking=True)
        for k, v in self.s_kps.items():
            v.copy_(kp_source[k], non_blocking=True)
        self.graph.replay()
        if check and self._strict:
            from .modules.dense_motion import check_status
            check_status(self.gen._eng.dm.status, clear=True)
        return self.out
