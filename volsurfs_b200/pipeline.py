"""The per-ray rendering hot path end to end: K-layer shell intersection -> packing -> appearance heads -> compositing
(forward and backward) — the B200 counterpart of ``VolSurfs.render_rays`` (volsurfs_py/methods/volsurfs.py:423-761) for the
mesh path with a constant background colour (volsurfs.py:686-689,708).

No host synchronisation happens between the stages: packing keeps capacity-sized (N*K) sample arrays and leaves the sample
count on the device (the reference syncs after every mesh trace and reads ``any_hit`` on the host, volsurfs.py:481)."""
from __future__ import annotations

import contextlib

import torch

from .appearance import AppearanceHead
from .raytracer import ShellTracer
from .volsurfs import VolumeRendering as VR


def ctypes_float3(t):
    """HOST float[3] for the entry points that take a background colour by value (built once: reading a device tensor synchronises)"""
    import ctypes

    return (ctypes.c_float * 3)(*[float(x) for x in t.detach().cpu().view(-1)[:3]])


@contextlib.contextmanager
def span(name: str):
    """NVTX range with the reference's profiler span names (mvdatasets/utils/profiler.py:4-103; spans ``meshes_raytracing``,
    ``ray_color_inference``, ``render_fg`` at volsurfs_py/methods/volsurfs.py:473-488,520-599,627-643 and ``forward_pass`` /
    ``backward_pass`` at callbacks/callback.py:79-107), so that nsys / ncu timelines of this path read like the reference's profile.
    Host-side markers only: nothing is synchronised (the reference's profiler relies on CUDA_LAUNCH_BLOCKING=1)."""
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


class ShellRenderer:
    def __init__(self, tracer: ShellTracer, rgb_head: AppearanceHead, alpha_head: AppearanceHead, bg_color=(1.0, 1.0, 1.0)):
        self.tracer = tracer
        self.rgb_head = rgb_head
        self.alpha_head = alpha_head
        self.K = tracer.nr_meshes
        self.bg_color = torch.tensor(bg_color, dtype=torch.float32, device=tracer.device)  # white (config/data_config.cfg:22-26)
        self._bg_host = ctypes_float3(torch.tensor(bg_color, dtype=torch.float32))
        self._loss_scratch = None  # 16 bytes the fused loss kernel accumulates in (self-clearing)
        self._zero_rows = None  # [N,1] zeros: the depth / accumulation outputs carry no gradient in a photometric loss

    # ---- stages ------------------------------------------------------------------------------------------------------
    def intersect_and_pack(self, rays_o, rays_d):
        """stage 1+2: all K layers in one launch, hits packed outer -> inner, face normals gathered"""
        with span("meshes_raytracing"):
            return self.tracer.render_samples(rays_o, rays_d, exact_size=False, with_normals=True)

    def shade_train(self, rsp, pos_features):
        """stage 3 in training mode: like ``shade`` but both heads keep their activations (per-head stash buffers owned by the
        renderer, re-used across steps) for ``heads_backward``"""
        S = int(pos_features.shape[0])
        outs = []
        with span("ray_color_inference"):
            for name, head in (("rgb", self.rgb_head), ("alpha", self.alpha_head)):
                stash = getattr(self, "_stash_" + name, None)
                if stash is None or stash.numel() < head.stash_bytes(S) or stash.device != pos_features.device:
                    stash = head.new_stash(S, pos_features.device)
                    setattr(self, "_stash_" + name, stash)
                out, _ = head.forward_train(pos_features, rsp.samples_dirs, rsp.samples_normals, n_valid_dev=rsp.total_dev, stash=stash)
                outs.append(out)
        return outs[0], outs[1]

    def shade(self, rsp, pos_features):
        """stage 3: per-hit colour and alpha (volsurfs.py:544-599).  ``pos_features`` [N*K, F]: output of the positional
        encoder for ``rsp.samples_3d`` (the permutohedral encoding is the stage before this path; synthetic in benchmarks)."""
        with torch.no_grad(), span("ray_color_inference"):  # forward and backward are driven explicitly (render_fwd_bwd); no autograd graph
            rgb = self.rgb_head(pos_features, rsp.samples_dirs, rsp.samples_normals, n_valid_dev=rsp.total_dev)
            alpha = self.alpha_head(pos_features, rsp.samples_dirs, rsp.samples_normals, n_valid_dev=rsp.total_dev)
        return rgb, alpha

    def composite(self, rsp, alpha, rgb):
        """stage 4 forward: (rgb_fg, depth, acc, bgT) and the composited prediction rgb_fg + bgT * bg (volsurfs.py:708)"""
        with span("render_fg"):
            rgb_fg, depth, acc, bgT = VR.composite(rsp, alpha, rgb, rsp.samples_z)
            pred = torch.addcmul(rgb_fg, bgT, self.bg_color.view(1, 3))
        return {"rgb": pred, "rgb_fg": rgb_fg, "depth": depth, "acc": acc, "bg_transmittance": bgT}

    def blend_l1(self, rgb_fg, bgT, gt_rgb):
        """background blend + L1 loss + the loss gradient in one launch (``vs_blend_l1_loss``: volsurfs.py:708, utils/losses.py:14-19
        and their autograd) -> (pred [N,3], loss 0-d, g_pred [N,3], g_bgT [N,1])"""
        from . import _lib

        with span("losses"):
            n = int(rgb_fg.shape[0])
            pred, g_pred = torch.empty_like(rgb_fg), torch.empty_like(rgb_fg)
            g_bgT, loss = torch.empty_like(bgT), torch.empty((), dtype=torch.float32, device=rgb_fg.device)
            if self._loss_scratch is None or self._loss_scratch.device != rgb_fg.device:
                self._loss_scratch = torch.zeros(2, dtype=torch.int64, device=rgb_fg.device)
            _lib.check(_lib.lib().vs_blend_l1_loss(rgb_fg.data_ptr(), bgT.data_ptr(), gt_rgb.contiguous().data_ptr(), self._bg_host,
                                                   pred.data_ptr(), g_pred.data_ptr(), g_bgT.data_ptr(), loss.data_ptr(),
                                                   self._loss_scratch.data_ptr(), n, torch.cuda.current_stream().cuda_stream),
                       "vs_blend_l1_loss")
        return pred, loss, g_pred, g_bgT

    def composite_l1(self, rsp, alpha, rgb, gt_rgb):
        """stage 4 forward of a TRAINING step: compositing, then ``blend_l1``.  Returns the dict of ``composite`` plus ``loss``,
        ``g_pred`` and ``g_bgT`` for ``composite_backward(..., g_bgT=...)``."""
        with span("render_fg"):
            rgb_fg, depth, acc, bgT = VR.composite(rsp, alpha, rgb, rsp.samples_z)
        pred, loss, g_pred, g_bgT = self.blend_l1(rgb_fg, bgT, gt_rgb)
        return {"rgb": pred, "rgb_fg": rgb_fg, "depth": depth, "acc": acc, "bg_transmittance": bgT, "loss": loss, "g_pred": g_pred,
                "g_bgT": g_bgT}

    def composite_backward(self, rsp, alpha, rgb, g_pred, g_bgT=None):
        """stage 4 backward for a loss on the composited prediction: g_rgb_fg = g_pred, g_bgT = g_pred . bg"""
        if g_bgT is None:
            g_bgT = (g_pred * self.bg_color.view(1, 3)).sum(dim=1, keepdim=True)
        if self._zero_rows is None or self._zero_rows.shape != g_bgT.shape or self._zero_rows.device != g_bgT.device:
            self._zero_rows = torch.zeros_like(g_bgT)
        zeros = self._zero_rows
        d_alpha, d_rgb, _ = VR.composite_backward(rsp, alpha, rgb, rsp.samples_z, g_pred.contiguous(), zeros, zeros, g_bgT)
        return d_alpha, d_rgb

    # ---- whole path --------------------------------------------------------------------------------------------------
    def render(self, rays_o, rays_d, pos_features):
        rsp = self.intersect_and_pack(rays_o, rays_d)
        rgb, alpha = self.shade(rsp, pos_features)
        out = self.composite(rsp, alpha, rgb)
        out.update(ray_samples_packed=rsp, samples_rgb=rgb, samples_alpha=alpha)
        return out

    def heads_backward(self, rsp, pos_features, d_rgb, d_alpha, want_feature_grads=True, fwd_outs=None):
        """stage 3 backward: per-sample colour / alpha gradients -> Linear parameter gradients of both heads (flat fp32 buffers
        ``grad_rgb`` / ``grad_alpha``, layout of AppearanceHead.split_flat) and, per head, the gradient of its positional features
        (every reference RGB model owns its encoder, rgb.py:40-60, so the two feature gradients stay separate)."""
        out = {}
        for name, head, g in (("rgb", self.rgb_head, d_rgb), ("alpha", self.alpha_head, d_alpha)):
            flat = getattr(self, "_grad_" + name, None)
            if flat is None or flat.device != pos_features.device:
                flat = torch.zeros(head.num_params(), dtype=torch.float32, device=pos_features.device)
                setattr(self, "_grad_" + name, flat)
            d_feat = None
            if want_feature_grads:
                d_feat = getattr(self, "_dfeat_" + name, None)
                if d_feat is None or d_feat.shape != pos_features.shape or d_feat.device != pos_features.device:
                    d_feat = torch.zeros_like(pos_features)
                    setattr(self, "_dfeat_" + name, d_feat)
            stash = getattr(self, "_stash_" + name, None) if fwd_outs is not None else None
            head.backward_into(pos_features, rsp.samples_dirs, rsp.samples_normals, g, flat, d_feat, False, rsp.total_dev,
                               stash=stash, fwd_out=None if stash is None else fwd_outs[name])
            out["grad_" + name] = flat
            out["d_features_" + name] = d_feat
        return out

    def render_fwd_bwd(self, rays_o, rays_d, pos_features, gt_rgb, heads_backward=True):
        """one pass of the hot path with the L1 photometric loss of the reference (utils/losses.py:14-19):
        forward, d loss / d pred, compositing backward down to per-sample colour and alpha gradients, then the backward of both
        appearance heads (parameter gradients + gradients of their positional features)"""
        with span("forward_pass"):
            if heads_backward:
                rsp = self.intersect_and_pack(rays_o, rays_d)
                rgb, alpha = self.shade_train(rsp, pos_features)
                out = self.composite_l1(rsp, alpha, rgb, gt_rgb)
                out.update(ray_samples_packed=rsp, samples_rgb=rgb, samples_alpha=alpha)
            else:
                out = self.render(rays_o, rays_d, pos_features)
                out["rgb"], out["loss"], out["g_pred"], out["g_bgT"] = self.blend_l1(out["rgb_fg"], out["bg_transmittance"], gt_rgb)
            loss, g_pred = out["loss"], out["g_pred"]
        with span("backward_pass"):
            rsp = out["ray_samples_packed"]
            d_alpha, d_rgb = self.composite_backward(rsp, out["samples_alpha"], out["samples_rgb"], g_pred, out["g_bgT"])
            out.update(loss=loss, d_alpha=d_alpha, d_rgb=d_rgb)
            if heads_backward:
                out.update(self.heads_backward(rsp, pos_features, d_rgb, d_alpha,
                                               fwd_outs={"rgb": out["samples_rgb"], "alpha": out["samples_alpha"]}))
        return out


class EncodedShellRenderer(ShellRenderer):
    """The legacy appearance with its real positional encoders: every head owns a permutohedral hash encoder (rgb.py:40-60,
    encodings/permutohash.py), so a training step is trace -> pack -> 2 x encode -> 2 x head -> composite -> loss -> composite backward
    -> 2 x head backward -> 2 x lattice scatter.  Like ShellRenderer it drives forward and backward explicitly, keeps capacity-sized
    buffers and the sample count on the device, and can therefore be captured into one CUDA graph (BASELINE config[3])."""

    def __init__(self, tracer: ShellTracer, rgb_head: AppearanceHead, alpha_head: AppearanceHead, rgb_encoder, alpha_encoder,
                 bg_color=(1.0, 1.0, 1.0)):
        super().__init__(tracer, rgb_head, alpha_head, bg_color)
        self.encoders = {"rgb": rgb_encoder, "alpha": alpha_encoder}
        self._bufs = {}

    def _buf(self, key, shape, zero=False):
        t = self._bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=self.tracer.device)
            self._bufs[key] = t
        return t

    def parameters(self):
        return ([p for e in self.encoders.values() for p in e.parameters()] + list(self.rgb_head.parameters())
                + list(self.alpha_head.parameters()))

    def render_fwd_bwd(self, rays_o, rays_d, gt_rgb, rgb_branch_only: bool = False):
        """one training step; returns loss, image and the gradients: ``grad_<head>`` (flat Linear gradients, AppearanceHead.split_flat
        layout) and ``grad_lattice_<head>`` (shaped like the encoder's lattice_values)"""
        rsp = self.intersect_and_pack(rays_o, rays_d)
        cap = rsp.get_max_nr_samples()
        heads = {"rgb": self.rgb_head, "alpha": self.alpha_head}
        feats, outs, stashes = {}, {}, {}
        for name, head in heads.items():
            enc = self.encoders[name]
            e = enc.encoder
            feats[name] = self._buf("feat_" + name, (cap, enc.output_dim), zero=True)
            e._launch_forward(e.lattice_values, rsp.samples_3d, enc.window(None), enc.output_dim, enc.bb_sides, rsp.total_dev, out=feats[name])
            stash = self._bufs.get("stash_" + name)
            if stash is None or stash.numel() < head.stash_bytes(cap):
                stash = self._bufs["stash_" + name] = head.new_stash(cap, self.tracer.device)
            stashes[name] = stash
            outs[name], _ = head.forward_train(feats[name], rsp.samples_dirs, rsp.samples_normals, n_valid_dev=rsp.total_dev, stash=stash,
                                               out=self._buf("out_" + name, (cap, head.out_dim)))
        comp = self.composite_l1(rsp, outs["alpha"], outs["rgb"], gt_rgb)
        d_alpha, d_rgb = self.composite_backward(rsp, outs["alpha"], outs["rgb"], comp["g_pred"], comp["g_bgT"])
        result = {"loss": comp["loss"], "rgb": comp["rgb"], "ray_samples_packed": rsp}
        self._state = (rsp, feats, outs, stashes, {"rgb": d_rgb, "alpha": d_alpha})
        for name in (("rgb",) if rgb_branch_only else ("rgb", "alpha")):
            result.update(self.branch_backward(name))
        return result

    def branch_backward(self, name):
        """backward of one head and its encoder from the state the forward part left behind: ``grad_<name>`` and ``grad_lattice_<name>``.
        ``render_fwd_bwd(..., rgb_branch_only=True)`` followed by ``branch_backward("alpha")`` lets a caller start the gradient exchange
        of the colour branch while the transparency branch is still computing."""
        rsp, feats, outs, stashes, d_out = self._state
        head = self.rgb_head if name == "rgb" else self.alpha_head
        enc = self.encoders[name]
        e = enc.encoder
        cap = rsp.get_max_nr_samples()
        flat = self._buf("grad_" + name, (head.num_params(),), zero=True)
        d_feat = self._buf("dfeat_" + name, (cap, enc.output_dim), zero=True)
        head.backward_into(feats[name], rsp.samples_dirs, rsp.samples_normals, d_out[name], flat, d_feat, False, rsp.total_dev,
                           stash=stashes[name], fwd_out=outs[name])
        d_lat = self._buf("dlat_" + name, tuple(e.lattice_values.shape), zero=True)
        d_lat.zero_()
        e._launch_backward(e.lattice_values, rsp.samples_3d, enc.window(None), d_feat, enc.bb_sides, rsp.total_dev, want_lattice=True,
                           d_lattice=d_lat, order_key=rsp.samples_layer)   # same-layer hits of neighbouring rays share lattice vertices
        return {"grad_" + name: flat, "grad_lattice_" + name: d_lat}


class GraphedTrainingStep:
    """``ShellRenderer.render_fwd_bwd`` captured once into a CUDA graph and replayed with one launch per step.

    The path has no host synchronisation (capacity-sized packed arrays, sample count on the device), so the ~45 kernels of a step
    can be replayed back to back; launched one by one from Python the device waits for the host between the short ones.  The
    tensors passed here are the graph's static inputs: copy new rays / features / targets INTO them before ``replay()``.  Results
    (``out`` of render_fwd_bwd, including the heads' flat gradient buffers) live in graph-owned memory and are overwritten by the
    next replay."""

    def __init__(self, renderer: "ShellRenderer", rays_o, rays_d, pos_features, gt_rgb, warmup: int = 2):
        self.renderer = renderer
        self.inputs = (rays_o, rays_d, pos_features, gt_rgb)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # allocates the renderer's cached buffers (stashes, workspaces, packed weights) outside the graph
            for _ in range(max(warmup, 1)):
                renderer.render_fwd_bwd(*self.inputs)
        torch.cuda.current_stream().wait_stream(side)
        assert not renderer.tracer.overflowed()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = renderer.render_fwd_bwd(*self.inputs)

    def replay(self):
        self.graph.replay()
        return self.out

    __call__ = replay


class PipelinedTrainingStep:
    """End-to-end training step fed from pinned HOST ray buffers, software-pipelined over the copy engines.

    Two CUDA graphs alternate.  Graph p runs ``render_fwd_bwd`` on device ray buffers p; forked inside the same graph, one side
    stream copies the NEXT step's rays (pinned host -> device buffers 1-p) and another copies the PREVIOUS step's image and loss
    (device buffers 1-p -> pinned host), so both transfers hide under the ~2.6 ms of compute instead of preceding / following it.
    Every step still moves one step's inputs in and one step's results out.

        loop = PipelinedTrainingStep(renderer, o_pin, d_pin, feats, gt, img_pin, loss_pin)
        loop.prime()                      # rays of step 0 -> device
        for i in range(steps):            # (write the rays of step i+1 into o_pin / d_pin before calling step(i))
            loop.step(i)                  # results of step i-1 are in img_pin / loss_pin once step(i) has completed
        loop.drain(steps)                 # results of the last step
    """

    def __init__(self, renderer: "ShellRenderer", o_pin, d_pin, pos_features, gt_rgb, img_pin, loss_pin, warmup: int = 2, step_fn=None):
        """``step_fn(rays_o, rays_d) -> dict`` (with "rgb" and "loss") replaces ``renderer.render_fwd_bwd(rays_o, rays_d, pos_features,
        gt_rgb)`` when given: any capturable step (e.g. one with real encoders and an in-graph gradient exchange) can be pipelined.
        ``img_pin = None``: only the loss comes back to the host every step (a training loop that does not look at the image)"""
        dev = gt_rgb.device
        if step_fn is None:
            step_fn = lambda o, d: renderer.render_fwd_bwd(o, d, pos_features, gt_rgb)  # noqa: E731
        self.renderer = renderer
        self.o_pin, self.d_pin, self.img_pin, self.loss_pin = o_pin, d_pin, img_pin, loss_pin
        self.rays_o = [torch.empty(o_pin.shape, dtype=torch.float32, device=dev) for _ in range(2)]
        self.rays_d = [torch.empty(d_pin.shape, dtype=torch.float32, device=dev) for _ in range(2)]
        self.img = [torch.empty(img_pin.shape, dtype=torch.float32, device=dev) for _ in range(2)] if img_pin is not None else None
        self.loss = [torch.empty((), dtype=torch.float32, device=dev) for _ in range(2)]
        self.copy_in, self.copy_out = torch.cuda.Stream(), torch.cuda.Stream()
        for b in range(2):
            self.rays_o[b].copy_(o_pin)
            self.rays_d[b].copy_(d_pin)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                step_fn(self.rays_o[0], self.rays_d[0])
        torch.cuda.current_stream().wait_stream(side)
        self.graphs, self.outs = [], []
        for b in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                main = torch.cuda.current_stream()
                self.copy_in.wait_stream(main)
                self.copy_out.wait_stream(main)
                with torch.cuda.stream(self.copy_in):
                    self.rays_o[1 - b].copy_(o_pin, non_blocking=True)
                    self.rays_d[1 - b].copy_(d_pin, non_blocking=True)
                with torch.cuda.stream(self.copy_out):
                    if img_pin is not None:
                        img_pin.copy_(self.img[1 - b], non_blocking=True)
                    loss_pin.copy_(self.loss[1 - b], non_blocking=True)
                out = step_fn(self.rays_o[b], self.rays_d[b])
                if img_pin is not None:
                    self.img[b].copy_(out["rgb"])
                self.loss[b].copy_(out["loss"])
                main.wait_stream(self.copy_in)
                main.wait_stream(self.copy_out)
            self.graphs.append(g)
            self.outs.append(out)

    def prime(self):
        self.rays_o[0].copy_(self.o_pin, non_blocking=True)
        self.rays_d[0].copy_(self.d_pin, non_blocking=True)

    def step(self, i: int):
        self.graphs[i & 1].replay()
        return self.outs[i & 1]

    def drain(self, n_steps: int):
        b = (n_steps - 1) & 1
        if self.img_pin is not None:
            self.img_pin.copy_(self.img[b], non_blocking=True)
        self.loss_pin.copy_(self.loss[b], non_blocking=True)


class TexturedShellRenderer:
    """The same path with the reference's DEFAULT appearance (config/volsurfs/base_5.cfg: using_neural_textures, per-surface independent
    colour and transparency models): every layer k owns an ``SHNeuralTextures`` for rgb (3 channels) and one for alpha (1 channel),
    queried with the hit's texture coordinates and the ray direction (volsurfs.py:535-599), alpha modulated by the no-grad decay of
    volsurfs.py:583-594, then composited.  Differentiable end to end through torch autograd: ``CompositeFunc`` (fused compositing
    backward) -> per-layer routing -> ``SHNeuralTextures`` (texture-network backward kernels).

    Hits are routed to their layer's models with one stable sort by layer per step (one host read of the per-layer counts; the
    reference reads ``any_hit`` on the host once per mesh)."""

    def __init__(self, tracer: ShellTracer, rgb_models, alpha_models, bg_color=(1.0, 1.0, 1.0), with_alpha_decay: bool = True,
                 n_streams: int = 4):
        assert len(rgb_models) == tracer.nr_meshes and len(alpha_models) == tracer.nr_meshes
        # the 2K models are independent and their kernels (narrow networks, atomics) leave most of an SM idle: spread them over a few
        # CUDA streams so that kernels of different models overlap; autograd replays every backward on its forward's stream
        self.streams = [torch.cuda.Stream(device=tracer.device) for _ in range(max(int(n_streams), 0))]
        self.tracer = tracer
        self.rgb_models = torch.nn.ModuleList(rgb_models)
        self.alpha_models = torch.nn.ModuleList(alpha_models)
        self.K = tracer.nr_meshes
        self.with_alpha_decay = bool(with_alpha_decay)
        self.bg_color = torch.tensor(bg_color, dtype=torch.float32, device=tracer.device)

    def parameters(self):
        return list(self.rgb_models.parameters()) + list(self.alpha_models.parameters())

    def intersect_and_pack(self, rays_o, rays_d):
        rsp = self.tracer.render_samples(rays_o, rays_d, exact_size=True, with_normals=True)
        rsp.samples_tex_uv = self.tracer.sample_uvs(rsp)
        return rsp

    def shade(self, rsp):
        """per-hit colour [S,3] and alpha [S,1]"""
        S = rsp.get_max_nr_samples()
        dev = rsp.samples_z.device
        if S == 0:
            return torch.zeros((0, 3), device=dev), torch.zeros((0, 1), device=dev)
        layer = rsp.samples_layer.view(-1).long()
        order = torch.argsort(layer, stable=True)
        counts = torch.bincount(layer, minlength=self.K).tolist()
        uv_s, dirs_s = rsp.samples_tex_uv[order], rsp.samples_dirs[order]
        rgb_parts, alpha_parts, o = [], [], 0
        main = torch.cuda.current_stream()
        jobs = []
        for k in range(self.K):
            n = counts[k]
            if n:
                jobs.append((self.rgb_models[k], rgb_parts, uv_s[o:o + n], dirs_s[o:o + n]))
                jobs.append((self.alpha_models[k], alpha_parts, uv_s[o:o + n], dirs_s[o:o + n]))
            o += n
        if self.streams:
            ready = torch.cuda.Event()
            ready.record(main)
            for i, (model, dst, uv_k, dirs_k) in enumerate(jobs):
                st = self.streams[i % len(self.streams)]
                st.wait_event(ready)
                with torch.cuda.stream(st):
                    out_k = model(uv_coords=uv_k, view_dirs=dirs_k)
                out_k.record_stream(main)
                dst.append(out_k)
            for st in self.streams:
                main.wait_stream(st)
        else:
            for model, dst, uv_k, dirs_k in jobs:
                dst.append(model(uv_coords=uv_k, view_dirs=dirs_k))
        inv = torch.empty_like(order)
        inv[order] = torch.arange(S, device=dev)
        rgb = torch.cat(rgb_parts)[inv]
        alpha = torch.cat(alpha_parts)[inv]
        if self.with_alpha_decay:
            with torch.no_grad():  # volsurfs.py:583-594
                dot = torch.sum(-rsp.samples_dirs * rsp.samples_normals, dim=1, keepdim=True).clamp(0.0, 1.0)
                decay = torch.sigmoid(10.0 * dot) * 2.0 - 1.0
            alpha = alpha * decay
        return rgb, alpha

    def render(self, rays_o, rays_d):
        from .volume_rendering import composite

        rsp = self.intersect_and_pack(rays_o, rays_d)
        rgb, alpha = self.shade(rsp)
        rgb_fg, depth, acc, bgT = composite(rsp, alpha, rgb)
        pred = rgb_fg + bgT * self.bg_color.view(1, 3)  # volsurfs.py:708
        return {"rgb": pred, "rgb_fg": rgb_fg, "depth": depth, "acc": acc, "bg_transmittance": bgT, "ray_samples_packed": rsp,
                "samples_rgb": rgb, "samples_alpha": alpha}

    def render_fwd_bwd(self, rays_o, rays_d, gt_rgb):
        """forward, L1 photometric loss (utils/losses.py:14-19), backward to every texture network's parameters (``.grad``)"""
        out = self.render(rays_o, rays_d)
        loss = (out["rgb"] - gt_rgb).abs().mean()
        loss.backward()
        out["loss"] = loss.detach()
        return out


def make_synthetic_textured_renderer(K=5, n_lat=224, n_lon=224, deg_res=(2048, 1024, 512, 256), sh_range=(15.0, 15.0, 15.0, 15.0), seed=0,
                                     offset=0.01, table_init=None):
    """C2-shaped scene with the default appearance: K nested shells with a lat/lon texture chart, per-layer SHNeuralTextures for colour
    and transparency configured like config/volsurfs/base_5.cfg (sh_degree 3, lerp, squeezing + 8-bit quantisation, align_to_webgl)"""
    from .synthetic import shell_face_uvs, shell_meshes
    from .textures import SHNeuralTextures

    meshes = shell_meshes(K=K, n_lat=n_lat, n_lon=n_lon, offset=offset)
    tracer = ShellTracer(meshes)
    tracer.set_face_uvs([shell_face_uvs(n_lat, n_lon)] * K)
    torch.manual_seed(4321 + seed)

    def model(c):
        m = SHNeuralTextures(sh_deg=3, nr_channels=c, sh_range=list(sh_range), anchor=False, lerp=True, deg_res=list(deg_res),
                             quantize_output=True, squeeze_output=True, align_to_webgl=True)
        if table_init is not None:  # tiny-cuda-nn starts at U(-1e-4, 1e-4): every texel equal; tests want some structure
            with torch.no_grad():
                for nt in m.neural_textures:
                    nt.model.table.copy_((torch.rand_like(nt.model.table) * 2 - 1) * table_init)
        return m.cuda()

    return TexturedShellRenderer(tracer, [model(3) for _ in range(K)], [model(1) for _ in range(K)]), meshes


def make_synthetic_renderer(K=5, n_lat=224, n_lon=224, hidden=(128, 128, 64), pos_dim=51, seed=0, offset=0.01, device=None):
    """C2/C5-shaped scene: K nested lumpy shells (~100k triangles each) + fixed-seed legacy heads (rgb: 3 outputs, alpha: 1 output
    with alpha decay, both normal-independent, GELU)"""
    from .synthetic import shell_meshes

    if device is not None:
        torch.cuda.set_device(device)
    meshes = shell_meshes(K=K, n_lat=n_lat, n_lon=n_lon, offset=offset)
    tracer = ShellTracer(meshes)
    torch.manual_seed(1234 + seed)
    rgb_head = AppearanceHead(pos_dim, hidden, 3, 3, False, "gelu", False).cuda()
    alpha_head = AppearanceHead(pos_dim, hidden, 1, 3, False, "gelu", True).cuda()
    return ShellRenderer(tracer, rgb_head, alpha_head), meshes
