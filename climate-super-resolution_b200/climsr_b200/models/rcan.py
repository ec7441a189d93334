"""Drop-in ``RCAN`` generator (the reference's default ``generator_type``, climsr/core/config.py:64) on the sm_100a kernels -
inference path.

Mirrors climsr/models/rcan.py:137-217: same constructor signature (``conv`` is accepted and must be the default), same
sub-module / parameter names (``head``, ``body.{g}.body.{b}.body.{0,2}`` + ``.body.3.conv_du.{0,2}``, ``tail.0.{0,2}``,
``tail.1``, ``srcnn.conv{1,2,3}``) and creation order - the parameter containers are plain nn.Conv2d - so
``load_state_dict`` (the reference's tolerant override included) and ``torch.manual_seed`` initialisation behave as before;
same ``forward(x, elev, mask) -> (N, 1, 4h, 4w)``.

Every 3x3 convolution runs on the tcgen05 conv kernel (csr_conv2d_nhwc) with bias / ReLU / the ResidualGroup and body
skip-adds fused in the epilogue; CALayer (global average pool -> 1x1 -> ReLU -> 1x1 -> sigmoid -> scale) together with the
RCAB skip is csr_channel_attention, PixelShuffle(2) is csr_pixel_shuffle2, the SRCNN tail runs on the same conv kernel.
Packed weights are cached per weight version.  Training (autograd) is not implemented for this model: forward under
``torch.enable_grad()`` with parameters that require gradients raises.  There is no CPU fallback.
"""
from __future__ import annotations

import logging
import math

import torch
import torch.nn as nn
from torch import Tensor

from .. import ops
from .._lib import CsrError, check, current_stream_ptr, lib


def default_conv(in_channels: int, out_channels: int, kernel_size: int, bias: bool = True) -> nn.Module:
    return nn.Conv2d(in_channels, out_channels, kernel_size, padding=kernel_size // 2, bias=bias)


class _CALayerParams(nn.Module):
    def __init__(self, channel: int, reduction: int = 16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv_du = nn.Sequential(nn.Conv2d(channel, channel // reduction, 1, padding=0, bias=True), nn.ReLU(inplace=True),
                                     nn.Conv2d(channel // reduction, channel, 1, padding=0, bias=True), nn.Sigmoid())


class _RCABParams(nn.Module):
    def __init__(self, n_feat: int, kernel_size: int, reduction: int):
        super().__init__()
        self.body = nn.Sequential(default_conv(n_feat, n_feat, kernel_size), nn.ReLU(True), default_conv(n_feat, n_feat, kernel_size),
                                  _CALayerParams(n_feat, reduction))
        self.res_scale = 1


class _ResidualGroupParams(nn.Module):
    def __init__(self, n_feat: int, kernel_size: int, reduction: int, n_resblocks: int):
        super().__init__()
        body = [_RCABParams(n_feat, kernel_size, reduction) for _ in range(n_resblocks)]
        body.append(default_conv(n_feat, n_feat, kernel_size))
        self.body = nn.Sequential(*body)


class _SRCNNParams(nn.Module):
    def __init__(self, in_channels: int = 3, out_channels: int = 1):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, kernel_size=9, padding=4)
        self.conv2 = nn.Conv2d(64, 32, kernel_size=1, padding=0)
        self.conv3 = nn.Conv2d(32, out_channels, kernel_size=5, padding=2)


class RCAN(nn.Module):
    def __init__(self, n_resgroups: int = 10, n_resblocks: int = 20, n_feats: int = 64, reduction: int = 16, scaling_factor: int = 4,
                 in_channels: int = 3, out_channels: int = 1, conv=default_conv, **kwargs):
        super().__init__()
        if conv is not default_conv and conv is not None:
            raise ValueError("climsr_b200 RCAN supports the reference's default_conv only")
        if n_feats != 64 or scaling_factor != 4 or out_channels != 1 or not (1 <= in_channels <= 16) or n_feats % reduction:
            raise ValueError("climsr_b200 RCAN supports n_feats=64, scaling_factor=4, out_channels=1, 1<=in_channels<=16")
        self.n_resgroups, self.n_resblocks, self.n_feats = n_resgroups, n_resblocks, n_feats
        self.kernel_size, self.reduction, self.scaling_factor = 3, reduction, scaling_factor
        self.in_channels = in_channels
        # creation order of rcan.py:164-186 (body, then head ... the reference builds the lists first, then the Sequentials in the
        # order head, body, tail, srcnn; parameters are created when the conv objects are): head conv, groups, body conv, tail
        head = [default_conv(in_channels, n_feats, 3)]
        body = [_ResidualGroupParams(n_feats, 3, reduction, n_resblocks) for _ in range(n_resgroups)]
        body.append(default_conv(n_feats, n_feats, 3))
        up = []
        for _ in range(int(math.log(scaling_factor, 2))):
            up += [default_conv(n_feats, 4 * n_feats, 3), nn.PixelShuffle(2)]
        tail = [nn.Sequential(*up), default_conv(n_feats, out_channels, 3)]
        self.head = nn.Sequential(*head)
        self.body = nn.Sequential(*body)
        self.tail = nn.Sequential(*tail)
        self.srcnn = _SRCNNParams(in_channels=3, out_channels=out_channels)

    # the reference's tolerant loader (rcan.py:189-217), kept verbatim in behaviour
    def load_state_dict(self, state_dict: dict, strict: bool = False) -> None:
        own_state = self.state_dict()
        for name, param in state_dict.items():
            if name in own_state:
                if isinstance(param, nn.Parameter):
                    param = param.data
                try:
                    own_state[name].copy_(param)
                except Exception:
                    if name.find("tail") >= 0:
                        logging.info("Replace pre-trained upsampler to new one...")
                    else:
                        raise RuntimeError(f"While copying the parameter named {name}, whose dimensions in the model are {own_state[name].size()} and "
                                           f"whose dimensions in the checkpoint are {param.size()}.")
            elif strict:
                if name.find("tail") == -1:
                    raise KeyError(f'unexpected key "{name}" in state_dict')
        if strict:
            missing = set(own_state.keys()) - set(state_dict.keys())
            if len(missing) > 0:
                raise KeyError(f'missing keys in state_dict: "{missing}"')

    # ------------------------------------------------------------------ packed-weight cache (see Discriminator._packed)
    def _packed(self, conv: nn.Conv2d):
        from .esrgan import _WEIGHT_EPOCH
        cache = self.__dict__.setdefault("_pack_cache", {})
        w, b = conv.weight, conv.bias
        key = (w.data_ptr(), w._version, b.data_ptr(), b._version, _WEIGHT_EPOCH[0], w.device)
        ent = cache.get(id(conv))
        if ent is not None and ent[1] == key:
            return ent[0], True
        kh, kw = conv.kernel_size
        nbytes = ops.conv2d_scratch_bytes(conv.out_channels, conv.in_channels, kh, kw)
        scratch = ent[0] if ent is not None and ent[0].numel() >= nbytes and ent[0].device == w.device else \
            torch.empty(max(nbytes, 16), dtype=torch.uint8, device=w.device)
        cache[id(conv)] = (scratch, key)
        return scratch, False

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_pack_cache", None)
        return state

    def __deepcopy__(self, memo):
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k != "_pack_cache":
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _conv(self, inp: Tensor, conv: nn.Conv2d, act: str = "none", res1=None, out_mode: str = "nhwc") -> Tensor:
        scratch, pre = self._packed(conv)
        return ops.conv2d_nhwc(inp, conv.weight.detach().contiguous().float(), conv.bias.detach().contiguous().float(), act=act, res1=res1,
                               out_mode=out_mode, scratch=scratch, prepacked=pre)

    # ------------------------------------------------------------------ forward (rcan.py:175-186)
    def forward(self, x: Tensor, elev: Tensor, mask: Tensor) -> Tensor:
        if x.dim() != 4 or elev.dim() != 4 or mask.dim() != 4:
            raise ValueError("expected x (N,C,h,w), elev (N,1,4h,4w), mask (N,1,4h,4w)")
        n, c, h, w = x.shape
        if c != self.in_channels:
            raise ValueError(f"x has {c} channels, RCAN was built with in_channels={self.in_channels}")
        hr_shape = (n, 1, 4 * h, 4 * w)
        if tuple(elev.shape) != hr_shape or tuple(mask.shape) != hr_shape:
            raise ValueError(f"elev/mask must have shape {hr_shape}, got {tuple(elev.shape)} / {tuple(mask.shape)}")
        if not x.is_cuda:
            raise CsrError("climsr_b200 RCAN runs on CUDA (sm_100a) only; there is no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise CsrError("climsr_b200 RCAN is inference-only (its backward is not implemented on the sm_100a path): call it under "
                           "torch.no_grad() or freeze its parameters")
        dev = x.device
        f32 = lambda t: t.detach().contiguous().float()  # noqa: E731
        with torch.cuda.device(dev):
            a = ops.nchw_to_nhwc_bf16(x.detach(), 64)
            head = self._conv(a, self.head[0])                                   # x = self.head(x)
            cur = head
            pooled = torch.empty((n, 64), dtype=torch.float32, device=dev)
            for grp in list(self.body)[:-1]:                                     # ResidualGroup (rcan.py:104-134)
                g_in = cur
                for blk in list(grp.body)[:-1]:                                  # RCAB (rcan.py:71-101)
                    t = self._conv(cur, blk.body[0], act="relu")
                    r = self._conv(t, blk.body[2])
                    ca = blk.body[3].conv_du
                    out = torch.empty_like(r)
                    check(lib.csr_channel_attention(r.data_ptr(), cur.data_ptr(), f32(ca[0].weight).data_ptr(), f32(ca[0].bias).data_ptr(),
                                                    f32(ca[2].weight).data_ptr(), f32(ca[2].bias).data_ptr(), out.data_ptr(), pooled.data_ptr(),
                                                    n, h, w, 64, ca[0].out_channels, current_stream_ptr()), "csr_channel_attention")
                    cur = out
                cur = self._conv(cur, grp.body[-1], res1=g_in)                   # res = body(x); res += x
            res = self._conv(cur, self.body[-1], res1=head)                      # res = self.body(x); res += x
            up = self.tail[0]
            t, hh, ww = res, h, w
            for i in range(0, len(up), 2):                                       # Upsampler: conv 64 -> 256, PixelShuffle(2)
                y4 = self._conv(t, up[i])
                t = torch.empty((n, 2 * hh, 2 * ww, 64), dtype=torch.bfloat16, device=dev)
                check(lib.csr_pixel_shuffle2(y4.data_ptr(), t.data_ptr(), n, hh, ww, 64, current_stream_ptr()), "csr_pixel_shuffle2")
                hh, ww = 2 * hh, 2 * ww
            y = self._conv(t, self.tail[1], out_mode="f32_planar")               # (N,1,H,W) fp32
            # x = self.srcnn(torch.cat([x, elev, mask], 1))   (srcnn.py:13-18)
            cat = torch.cat([y, elev.detach().to(dev).float(), mask.detach().to(dev).float()], 1).contiguous()
            s_in = ops.nchw_to_nhwc_bf16(cat, 64)
            s1 = self._conv(s_in, self.srcnn.conv1, act="relu")
            s2 = self._conv(s1, self.srcnn.conv2, act="relu")
            return self._conv(s2, self.srcnn.conv3, out_mode="f32_planar")
