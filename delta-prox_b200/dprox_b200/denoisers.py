"""Denoisers behind `deep_prior`.

`FFDNetColor` has the architecture and `state_dict` keys of the reference's FFDNet-color
(pnp/denoisers/models/network_ffdnet.py:27-68: PixelUnshuffle(2) -> conv3x3(13->96)+ReLU ->
10 x [conv3x3(96->96)+ReLU] -> conv3x3(96->12) -> PixelShuffle(2)), so the published
`ffdnet_color.pth` loads unchanged.  It is launched as an *external* prox between two native
stages on the same CUDA stream.  `precision="bf16"` runs the hand-written tcgen05 kernels of
csrc/dpx_conv_tc.cuh (forward and data gradient); `precision="fp32"` keeps fp32 library
convolutions for 1e-5-class parity with the reference.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .proxfn import Denoiser


class FFDNet(nn.Module):
    def __init__(self, in_nc=3, out_nc=3, nc=96, nb=12):
        super().__init__()
        layers = [nn.Conv2d(in_nc * 4 + 1, nc, 3, padding=1), nn.ReLU(inplace=True)]
        for _ in range(nb - 2):
            layers += [nn.Conv2d(nc, nc, 3, padding=1), nn.ReLU(inplace=True)]
        layers += [nn.Conv2d(nc, out_nc * 4, 3, padding=1)]
        self.model = nn.Sequential(*layers)          # keys model.{0,2,...,22}.{weight,bias}

    def forward(self, x, sigma):
        h, w = x.shape[-2:]
        x = F.pad(x, (0, (-w) % 2, 0, (-h) % 2), mode="replicate")
        x = F.pixel_unshuffle(x, 2)
        m = sigma.reshape(-1, 1, 1, 1).to(x.dtype).expand(x.shape[0], 1, x.shape[2], x.shape[3])
        x = self.model(torch.cat((x, m), 1))
        return F.pixel_shuffle(x, 2)[..., :h, :w]


class NativeFFDNet:
    """FFDNet-color on tcgen05 tensor cores through the C-ABI (`dpx_ffdnet_*`, csrc/dpx_conv_tc.cuh): channel-group-major
    bf16 activations, hand-written implicit-GEMM 3x3 convolutions (CTA pairs, TMEM accumulators, TMA-staged rows reused for all
    nine taps, resident filter bank), fused bias+ReLU, fused unshuffle/sigma prologue and shuffle/crop epilogue; the data
    gradient runs on the same kernel.  bf16 operands: ~1e-2 relative to the fp32 network; `split=True`: fp16 operand pairs
    (hi + 2^-11 lo'), ~1e-6 relative to the fp32 network at three times the tensor-core work."""

    def __init__(self, model: "FFDNet", device, split: bool = False):
        import ctypes as C
        from . import _cabi as cabi
        self._cabi, self.device = cabi, torch.device(device)
        self.profile = None            # set to a list to collect (kind, start event, end event, pixels) per call (bench.py)
        lib = cabi.lib()
        convs = [m for m in model.model if isinstance(m, nn.Conv2d)]
        self.n_layers = len(convs)
        self._h = C.c_void_p()
        cabi.check(lib.dpx_ffdnet_create(len(convs), convs[1].out_channels, C.byref(self._h)), "dpx_ffdnet_create")
        # split = fp16 hi + 2^-11 lo' operand pairs, 3 MMAs per k-step: fp32-class accuracy on the same tensor-core kernel
        self.split = bool(split)
        cabi.check(lib.dpx_ffdnet_set_precision(self._h, int(self.split)), "dpx_ffdnet_set_precision")
        with torch.cuda.device(self.device):
            for i, c in enumerate(convs):
                w = c.weight.detach().to(self.device, torch.float32).contiguous()
                b = c.bias.detach().to(self.device, torch.float32).contiguous()
                cabi.check(lib.dpx_ffdnet_set_layer(self._h, i, cabi.ptr(w), cabi.ptr(b), w.shape[0], w.shape[1],
                                                    cabi.stream_ptr(self.device)), "dpx_ffdnet_set_layer")
            torch.cuda.current_stream(self.device).synchronize()

    def _args(self, x, sigma):
        cabi = self._cabi
        x = cabi.require_cuda_f32(x, "x")
        B, Cc, H, W = x.shape
        if Cc != 3:
            raise ValueError("native FFDNet-color expects 3 channels")
        sigma = cabi.require_cuda_f32(sigma.to(x.device, torch.float32).reshape(-1), "sigma")
        if sigma.numel() not in (1, B):
            raise ValueError(f"sigma must have 1 or {B} entries")
        return x, sigma, B, H, W

    def __call__(self, x: torch.Tensor, sigma: torch.Tensor, train: bool = False) -> torch.Tensor:
        cabi = self._cabi
        x, sigma, B, H, W = self._args(x, sigma)
        y = torch.empty_like(x)
        fn = cabi.lib().dpx_ffdnet_forward_train if train else cabi.lib().dpx_ffdnet_forward
        ev = self._mark()
        with torch.cuda.device(x.device):
            cabi.check(fn(self._h, cabi.ptr(x), cabi.ptr(sigma), int(sigma.numel() > 1), cabi.ptr(y), B, H, W,
                          cabi.stream_ptr(x.device)), "dpx_ffdnet_forward")
        self._mark(ev, "fwd", B * H * W)
        return y

    def _mark(self, start=None, kind=None, pixels=0):
        if self.profile is None:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        if start is not None:
            self.profile.append((kind, start, e, pixels))
        return e

    def backward(self, g_y: torch.Tensor, n_sigma: int):
        """(dL/dx, dL/dsigma) of the last `train=True` call."""
        cabi = self._cabi
        g_y = cabi.require_cuda_f32(g_y, "g_y")
        B, _, H, W = g_y.shape
        g_x = torch.empty_like(g_y)
        g_s = torch.empty(n_sigma, device=g_y.device, dtype=torch.float32)
        ev = self._mark()
        with torch.cuda.device(g_y.device):
            cabi.check(cabi.lib().dpx_ffdnet_backward(self._h, cabi.ptr(g_y), cabi.ptr(g_x), cabi.ptr(g_s), int(n_sigma > 1), B, H, W,
                                                      cabi.stream_ptr(g_y.device)), "dpx_ffdnet_backward")
        self._mark(ev, "bwd", B * H * W)
        return g_x, g_s

    def load_weights(self, convs):
        """(re)pack the filter banks from the modules' current weights (after an optimizer step)"""
        cabi, lib = self._cabi, self._cabi.lib()
        with torch.cuda.device(self.device):
            for i, c in enumerate(convs):
                w = c.weight.detach().to(self.device, torch.float32).contiguous()
                b = c.bias.detach().to(self.device, torch.float32).contiguous()
                cabi.check(lib.dpx_ffdnet_set_layer(self._h, i, cabi.ptr(w), cabi.ptr(b), w.shape[0], w.shape[1],
                                                    cabi.stream_ptr(self.device)), "dpx_ffdnet_set_layer")

    def backward_params(self, g_y: torch.Tensor, n_sigma: int, shapes):
        """(dL/dx, dL/dsigma, [dL/dW_l], [dL/db_l]) of the last `train=True` call: data gradient and weight gradient both on the
        tensor-core kernels (dpx_ffdnet_backward_params)."""
        import ctypes as C
        cabi = self._cabi
        g_y = cabi.require_cuda_f32(g_y, "g_y")
        B, _, H, W = g_y.shape
        g_x = torch.empty_like(g_y)
        g_s = torch.empty(n_sigma, device=g_y.device, dtype=torch.float32)
        gws = [torch.empty(sw, device=g_y.device, dtype=torch.float32) for sw, _ in shapes]
        gbs = [torch.empty(sb, device=g_y.device, dtype=torch.float32) for _, sb in shapes]
        pw = (C.c_void_p * len(gws))(*[t.data_ptr() for t in gws])
        pb = (C.c_void_p * len(gbs))(*[t.data_ptr() for t in gbs])
        ev = self._mark()
        with torch.cuda.device(g_y.device):
            cabi.check(cabi.lib().dpx_ffdnet_backward_params(self._h, cabi.ptr(g_y), cabi.ptr(g_x), cabi.ptr(g_s), int(n_sigma > 1), pw, pb,
                                                             B, H, W, cabi.stream_ptr(g_y.device)), "dpx_ffdnet_backward_params")
        self._mark(ev, "bwd", B * H * W)
        return g_x, g_s, gws, gbs

    def conv_layer(self, layer: int, x: torch.Tensor, direction: int = 0, relu: bool = False) -> torch.Tensor:
        """one convolution of the network on fp32 NCHW tensors (per-layer parity tests)"""
        cabi = self._cabi
        x = cabi.require_cuda_f32(x, "x")
        B, Cc, H, W = x.shape
        head, tail = layer == 0, layer == self.n_layers - 1
        cin, cout = (13 if head else 96), (12 if tail else 96)
        if direction:
            cin, cout = cout, cin
        if Cc != cin:
            raise ValueError(f"layer {layer} expects {cin} channels in this direction, got {Cc}")
        y = torch.empty(B, cout, H, W, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            cabi.check(cabi.lib().dpx_ffdnet_conv_layer(self._h, layer, int(direction), int(relu), cabi.ptr(x), cabi.ptr(y), B, H, W,
                                                        cabi.stream_ptr(x.device)), "dpx_ffdnet_conv_layer")
        return y

    def wgrad_layer(self, layer: int, x: torch.Tensor, gy: torch.Tensor):
        """(dL/dW, dL/db) of one convolution given its input `x` and the gradient `gy` w.r.t. its pre-activation output, on the
        tensor-core weight-gradient kernel (csrc/dpx_conv_wgrad.cuh; bf16 operands)"""
        cabi = self._cabi
        x, gy = cabi.require_cuda_f32(x, "x"), cabi.require_cuda_f32(gy, "gy")
        B, cin, H, W = x.shape
        cout = gy.shape[1]
        gw = torch.empty(cout, cin, 3, 3, device=x.device, dtype=torch.float32)
        gb = torch.empty(cout, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            cabi.check(cabi.lib().dpx_ffdnet_wgrad_layer(self._h, int(layer), cabi.ptr(x), cabi.ptr(gy), cabi.ptr(gw), cabi.ptr(gb), B, H, W,
                                                         cabi.stream_ptr(x.device)), "dpx_ffdnet_wgrad_layer")
        return gw, gb

    def __deepcopy__(self, memo):
        raise TypeError("NativeFFDNet owns device filter banks; copy the owning denoiser instead (it rebuilds them lazily)")

    def __del__(self):
        try:
            if self._h:
                self._cabi.lib().dpx_ffdnet_destroy(self._h)
        except Exception:
            pass


class _NativeFFDNetFn(torch.autograd.Function):
    """y = FFDNet(x, sigma) with frozen weights: forward and data gradient on the tensor-core kernels."""
    _calls = 0

    @staticmethod
    def forward(ctx, x, sigma, net):
        ctx.net, ctx.sigma_shape = net, sigma.shape
        ctx.n_sigma = int(sigma.numel())
        _NativeFFDNetFn._calls += 1                                # process-wide, never reused
        ctx.graph_id = net._graph_id = _NativeFFDNetFn._calls
        ctx.save_for_backward(x.detach(), sigma.detach())
        return net(x.detach(), sigma.detach(), train=True)

    @staticmethod
    def backward(ctx, g):
        net = ctx.net
        if ctx.graph_id != net._graph_id:
            # another forward ran in between (an unrolled solver calls the denoiser once per iteration and backpropagates
            # in reverse order): the saved activations were overwritten -> recompute this call's forward
            x, sigma = ctx.saved_tensors
            net(x, sigma, train=True)
            net._graph_id = ctx.graph_id
        g_x, g_s = net.backward(g.contiguous(), ctx.n_sigma)
        net._graph_id = -1                                        # the saved state is consumed
        return g_x, g_s.reshape(ctx.sigma_shape), None


class _NativeFFDNetTrainFn(torch.autograd.Function):
    """y = FFDNet(x, sigma; weights) with TRAINABLE weights: forward, data gradient and weight gradient on the tensor-core kernels.
    `params` = (w_0, b_0, w_1, b_1, ...) of the convolutions, passed so that autograd routes their gradients."""

    @staticmethod
    def forward(ctx, x, sigma, net, convs, *params):
        ctx.net, ctx.convs, ctx.sigma_shape = net, convs, sigma.shape
        ctx.n_sigma = int(sigma.numel())
        ctx.shapes = [(tuple(c.weight.shape), tuple(c.bias.shape)) for c in convs]
        net.load_weights(convs)                                    # the banks follow the optimizer
        _NativeFFDNetFn._calls += 1
        ctx.graph_id = net._graph_id = _NativeFFDNetFn._calls
        ctx.save_for_backward(x.detach(), sigma.detach())
        return net(x.detach(), sigma.detach(), train=True)

    @staticmethod
    def backward(ctx, g):
        net = ctx.net
        if ctx.graph_id != net._graph_id:                          # another forward ran in between: recompute this call's
            x, sigma = ctx.saved_tensors
            net(x, sigma, train=True)
            net._graph_id = ctx.graph_id
        g_x, g_s, gws, gbs = net.backward_params(g.contiguous(), ctx.n_sigma, ctx.shapes)
        net._graph_id = -1
        grads = []
        for gw_, gb_ in zip(gws, gbs):
            grads += [gw_, gb_]
        return (g_x, g_s.reshape(ctx.sigma_shape), None, None, *grads)


class FFDNetColorDenoiser(Denoiser):
    """pnp/denoisers/wrapper.py:38-48.  `precision='fp32'` (default): the native tcgen05 network with fp16 operand pairs
    (fp32-class accuracy, 1e-5 parity with the reference's fp32 path); `precision='bf16'`: the same kernels with bf16 operands
    (the fast mode); `precision='torch'`: the framework's fp32 convolutions (comparison only)."""

    def __init__(self, model_path=None, seed=None, precision="fp32"):
        super().__init__()
        if precision not in ("fp32", "bf16", "torch"):
            raise ValueError("precision must be 'fp32', 'bf16' or 'torch'")
        self.precision, self._native = precision, None
        self.model = FFDNet(3, 3, 96, 12)
        if model_path is not None:
            self.model.load_state_dict(torch.load(model_path, map_location="cpu"), strict=True)
        elif seed is not None:
            self.load_seeded(seed)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_native" else copy.deepcopy(v, memo)      # filter banks are rebuilt lazily
        return new

    def load_seeded(self, seed: int):
        """Deterministic random weights (nn.Conv2d's default init bounds) — what the parity tests use, because the
        pretrained file needs a download (pnp/prior.py:14-35)."""
        g = torch.Generator().manual_seed(seed)
        convs = [m for m in self.model.model if isinstance(m, nn.Conv2d)]
        with torch.no_grad():
            for c in convs:
                bound = 1.0 / math.sqrt(c.in_channels * 9)
                c.weight.copy_((torch.rand(c.weight.shape, generator=g) * 2 - 1) * bound)
                c.bias.copy_((torch.rand(c.bias.shape, generator=g) * 2 - 1) * bound)
        return self

    def _native_net(self, device):
        if self._native is None or self._native.device != device:
            self._native = NativeFFDNet(self.model, device, split=self.precision == "fp32")
        return self._native

    def _denoise(self, x, sigma):
        trainable = any(p.requires_grad for p in self.model.parameters())
        wants_grad = torch.is_grad_enabled() and (x.requires_grad or sigma.requires_grad or trainable)
        native = self.precision in ("fp32", "bf16")
        if native and wants_grad and not trainable:
            # unrolled training with a frozen denoiser (BASELINE config 5, e2e_optics_dprox.py:34): forward and data gradient
            # both run on the native tensor-core kernels (activation recomputation when several calls are in flight)
            return _NativeFFDNetFn.apply(x.contiguous(), sigma, self._native_net(x.device))
        if native and not wants_grad:
            return self._native_net(x.device)(x, sigma)
        if self.precision == "bf16":
            # trainable denoiser weights: forward, data gradient AND weight gradient on the native tensor-core kernels
            convs = [m for m in self.model.model if isinstance(m, nn.Conv2d)]
            params = []
            for c in convs:
                params += [c.weight, c.bias]
            return _NativeFFDNetTrainFn.apply(x.contiguous(), sigma, self._native_net(x.device), convs, *params)
        # framework fp32 convolutions (no TF32): trainable weights at fp32, or precision='torch'
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            return self.model(x, sigma)
        finally:
            torch.backends.cudnn.allow_tf32 = prev


# ------------------------------------------------------------------------------------------------
#  DRUNet (pnp/denoisers/models/network_unet.py:67-116, wrapper.py:89-146) — SURVEY §8f rank 4
# ------------------------------------------------------------------------------------------------

class _ResBlock(nn.Module):
    """x + conv(relu(conv(x))), bias-free 3x3 convolutions; state_dict keys `res.0.weight`, `res.2.weight`."""

    def __init__(self, c):
        super().__init__()
        self.res = nn.Sequential(nn.Conv2d(c, c, 3, padding=1, bias=False), nn.ReLU(inplace=True), nn.Conv2d(c, c, 3, padding=1, bias=False))

    def forward(self, x):
        return x + self.res(x)


class UNetRes(nn.Module):
    """DRUNet body: head conv, 3 x (nb residual blocks + 2x2 stride-2 conv), nb residual blocks, 3 x (2x2 transposed conv +
    nb residual blocks) with additive skips, tail conv.  Module names follow the reference so published weights load."""

    def __init__(self, in_nc=2, out_nc=1, nc=(64, 128, 256, 512), nb=4):
        super().__init__()
        blocks = lambda c: [_ResBlock(c) for _ in range(nb)]
        self.m_head = nn.Conv2d(in_nc, nc[0], 3, padding=1, bias=False)
        self.m_down1 = nn.Sequential(*blocks(nc[0]), nn.Conv2d(nc[0], nc[1], 2, stride=2, bias=False))
        self.m_down2 = nn.Sequential(*blocks(nc[1]), nn.Conv2d(nc[1], nc[2], 2, stride=2, bias=False))
        self.m_down3 = nn.Sequential(*blocks(nc[2]), nn.Conv2d(nc[2], nc[3], 2, stride=2, bias=False))
        self.m_body = nn.Sequential(*blocks(nc[3]))
        self.m_up3 = nn.Sequential(nn.ConvTranspose2d(nc[3], nc[2], 2, stride=2, bias=False), *blocks(nc[2]))
        self.m_up2 = nn.Sequential(nn.ConvTranspose2d(nc[2], nc[1], 2, stride=2, bias=False), *blocks(nc[1]))
        self.m_up1 = nn.Sequential(nn.ConvTranspose2d(nc[1], nc[0], 2, stride=2, bias=False), *blocks(nc[0]))
        self.m_tail = nn.Conv2d(nc[0], out_nc, 3, padding=1, bias=False)

    def forward(self, x0):
        x1 = self.m_head(x0)
        x2 = self.m_down1(x1)
        x3 = self.m_down2(x2)
        x4 = self.m_down3(x3)
        x = self.m_body(x4)
        x = self.m_up3(x + x4)
        x = self.m_up2(x + x3)
        x = self.m_up1(x + x2)
        return self.m_tail(x + x1)


class DRUNetDenoiser(Denoiser):
    """DRUNet behind `deep_prior(x, denoiser='drunet' | 'drunet_color')` (wrapper.py:89-146): the noise level enters as an
    extra input channel; images larger than 256 x 256 are denoised as four overlapping quadrants (recursively), images up
    to that size are replicate-padded to a multiple of 16.  An opaque torch module on the prox hook, like every denoiser
    but FFDNet-color."""

    REFIELD, MIN_SIZE, MODULO = 32, 256, 16

    def __init__(self, n_channels=1, model_path=None):
        super().__init__()
        self.model = UNetRes(in_nc=n_channels + 1, out_nc=n_channels)
        if model_path is not None:
            self.model.load_state_dict(torch.load(model_path, map_location="cpu"), strict=True)

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k == "_native" else copy.deepcopy(v, memo)      # filter banks are rebuilt lazily
        return new

    def load_seeded(self, seed: int):
        """Deterministic random weights in state_dict order (U(-b, b), b = 1/sqrt(fan_in)) for the parity tests."""
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for _, w in self.model.state_dict().items():
                bound = 1.0 / math.sqrt(w[0].numel())
                w.copy_((torch.rand(w.shape, generator=g) * 2 - 1) * bound)
        return self

    def _denoise(self, x, sigma):
        if sigma.shape[0] != x.shape[0]:
            sigma = sigma.expand(x.shape[0], 1, 1, 1)
        inp = torch.cat((x, sigma.to(x.dtype).expand(x.shape[0], 1, x.shape[2], x.shape[3])), dim=1)
        prev = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            return self._tiled(inp)
        finally:
            torch.backends.cudnn.allow_tf32 = prev

    def _tiled(self, L):
        h, w = L.shape[-2:]
        if h * w <= self.MIN_SIZE ** 2:
            ph, pw = (-h) % self.MODULO, (-w) % self.MODULO
            return self.model(F.pad(L, (0, pw, 0, ph), mode="replicate"))[..., :h, :w]
        th, tw = (h // 2 // self.REFIELD + 1) * self.REFIELD, (w // 2 // self.REFIELD + 1) * self.REFIELD
        rows, cols = (slice(0, th), slice(h - th, h)), (slice(0, tw), slice(w - tw, w))
        run = self.model if h * w <= 4 * self.MIN_SIZE ** 2 else self._tiled
        E = None
        for i, rs in enumerate(rows):
            for j, cs in enumerate(cols):
                e = run(L[..., rs, cs])
                if E is None:
                    E = torch.zeros(e.shape[0], e.shape[1], h, w, dtype=L.dtype, device=L.device)
                # each quadrant of the output takes the matching corner of its (larger, overlapping) tile
                dst_r = slice(0, h // 2) if i == 0 else slice(h // 2, h)
                dst_c = slice(0, w // 2) if j == 0 else slice(w // 2, w)
                src_r = slice(0, h // 2) if i == 0 else slice(th - (h - h // 2), th)
                src_c = slice(0, w // 2) if j == 0 else slice(tw - (w - w // 2), tw)
                E[..., dst_r, dst_c] = e[..., src_r, src_c]
        return E


def get_denoiser(name: str):
    """pnp/prior.py:14-35.  Weight files are looked up under $DPROX_WEIGHTS (no network access here)."""
    root = os.environ.get("DPROX_WEIGHTS", os.path.expanduser("~/.cache/dprox/pnp_denoisers"))
    if name == "ffdnet_color":
        path = os.path.join(root, "ffdnet_color.pth")
        if not os.path.exists(path):
            raise FileNotFoundError(f"pretrained weights for {name!r} not found at {path}; set $DPROX_WEIGHTS or pass a "
                                    f"Denoiser instance to deep_prior(x, denoiser=obj)")
        return FFDNetColorDenoiser(path)
    if name in ("drunet", "drunet_color"):
        path = os.path.join(root, "drunet_gray.pth" if name == "drunet" else "drunet_color.pth")
        if not os.path.exists(path):
            raise FileNotFoundError(f"pretrained weights for {name!r} not found at {path}; set $DPROX_WEIGHTS or pass a "
                                    f"Denoiser instance to deep_prior(x, denoiser=obj)")
        return DRUNetDenoiser(1 if name == "drunet" else 3, path)
    raise NotImplementedError(f"denoiser {name!r} is not part of the lowered path (SURVEY §2 row 12); pass a Denoiser object")
