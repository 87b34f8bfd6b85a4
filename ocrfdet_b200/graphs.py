"""CUDA-graph replay of a whole render step (scope row f-2: the batched call site of view_transformer_ocrf.py:1090).

In capacity mode (`render_batch(pair_capacity=...)`) a step has no host synchronisation: preprocess, binning, both
blend directions and the preprocess backward are ~17 launches of fixed grid sizes, every data-dependent count lives on
the device.  `GraphedRenderStep` captures forward AND backward for fixed shapes once and replays them with one launch;
the programmatic-dependent-launch edges between the chain kernels are part of the captured graph.  Host issue time
drops from ~0.24 ms to the cost of one graph launch, which matters when the surrounding training step is launch bound
(the reference issues ~12 kernels and a blocking read-back per VIEW).

Usage:
    step = GraphedRenderStep(S=1, P=100_000, cams=cams, height=256, width=704, channels=3, pair_capacity=5_000_000)
    outs, grads = step(means3D=..., scales=..., rotations=..., opacities=..., colors=..., grad_color=..., grad_opacity=...)
    step.check_overflow()     # after any number of replays: raises if a replay needed more than pair_capacity pairs
`pre` / `post` hooks are captured with the step (host <-> device copies of pinned buffers, a loss): `step.graph.replay()`
is then a complete host-to-host step; two instances replayed on two streams overlap the copies of one step with the
kernels of the other (bench.py's `e2e`).
"""
import torch

from . import rasterizer as R

INPUTS = ("means3D", "scales", "rotations", "opacities", "colors")


class GraphedRenderStep:
    def __init__(self, S, P, cams, height, width, channels=3, pair_capacity=None, bg=None, binning=None, device="cuda",
                 pre=None, post=None):
        if pair_capacity is None:
            raise ValueError("a captured step needs a fixed pair_capacity (exact sizing reads the pair count on the host)")
        self.S, self.P, self.H, self.W, self.C = int(S), int(P), int(height), int(width), int(channels)
        self.cams = cams.to(device).float().contiguous()
        self.V = self.cams.shape[0]
        self.capacity, self.binning = int(pair_capacity), binning
        dev = torch.device(device)
        self.bg = (torch.zeros(self.C, device=dev) if bg is None else bg.to(dev).float()).contiguous()
        shapes = {"means3D": (S, P, 3), "scales": (S, P, 3), "rotations": (S, P, 4), "opacities": (S, P, 1),
                  "colors": (S, P, self.C)}
        self.static_in = {k: torch.zeros(v, device=dev).requires_grad_(True) for k, v in shapes.items()}
        self.static_gcolor = torch.zeros(self.V, self.C, self.H, self.W, device=dev)
        self.static_gopac = torch.zeros(self.V, 1, self.H, self.W, device=dev)
        # optional hooks captured WITH the step: pre(step) runs before the forward (e.g. non-blocking copies of pinned
        # host parameters into `static_in`), post(step, outs, grads) after the backward (e.g. the loss and non-blocking
        # copies of the gradients into pinned host buffers): a whole host-to-host training step is then ONE launch
        self.pre, self.post = pre, post
        self.graph = None
        self.outs = None
        self.grads = None

    def _step(self):
        t = self.static_in
        if self.pre is not None:
            with torch.no_grad():
                self.pre(self)
        color, radii, depth, opac = R.render_batch(t["means3D"], t["opacities"], self.cams, self.H, self.W, self.bg,
                                                   colors_precomp=t["colors"], scales=t["scales"],
                                                   rotations=t["rotations"], pair_capacity=self.capacity,
                                                   binning=self.binning)
        grads = torch.autograd.grad([color, opac], [t[k] for k in INPUTS], [self.static_gcolor, self.static_gopac])
        outs, grads = (color, radii, depth, opac), dict(zip(INPUTS, grads))
        if self.post is not None:
            with torch.no_grad():
                self.post(self, outs, grads)
        return outs, grads

    def capture(self, **example):
        """Fill the static buffers with `example` (representative inputs: they only serve as warm-up), run the step a
        few times on a side stream, then capture it."""
        self._load(example)
        status = R._Status.get(self.cams.device)
        status.muted = True  # an overflow of the warm-up inputs must not raise out of the middle of the capture;
        try:                 # it stays in the device's sticky words and is reported by check_overflow()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outs, self.grads = self._step()
        finally:
            status.muted = False
        return self

    def _load(self, inputs):
        with torch.no_grad():
            for k in INPUTS:
                if k in inputs and inputs[k] is not None:
                    self.static_in[k].copy_(inputs[k].reshape(self.static_in[k].shape))
            if inputs.get("grad_color") is not None:
                self.static_gcolor.copy_(inputs["grad_color"])
            if inputs.get("grad_opacity") is not None:
                self.static_gopac.copy_(inputs["grad_opacity"])

    def __call__(self, **inputs):
        """Copy the given tensors into the static buffers, replay, and return (color, radii, depth, opacity), grads.
        The returned tensors are the graph's static outputs: they are overwritten by the next replay."""
        if self.graph is None:
            self.capture(**inputs)
        self._load(inputs)
        self.graph.replay()
        return self.outs, self.grads

    def check_overflow(self):
        """Raise if ANY replay since the last check overflowed `pair_capacity`.  Every replay ORs its error flags into
        the device's sticky status words (`sticky_status` of ocrf_bin_forward), which no forward clears -- the geom
        header itself is re-zeroed by each replay's preprocess -- so one read-back covers any number of replays.
        The captured step also mirrors those words into pinned host memory, so the same error surfaces without a
        synchronisation at the next eager `render_batch` call."""
        R.check_overflow(self.cams.device)
