"""Feature stack + correspondence on the B200 path: host-side mirror of the two consumers of `extract` that the hot
path names - aggregation_network.py:62-66 (bilinear resize to 128x128 + channel concat) and
correspondence_utils.py:113-146 (find_nn_source_correspondences / points_to_idxs)."""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from ._lib import check, ptr, stream_ptr


def build_stack(feats, out_hw=(128, 128), layout="nhwc"):
    """feats: dict id -> fp16 (B, C, h, w) token-major views as returned by FeatureExtractor.extract (or a list of
    them). Returns the stack of aggregation_network.py:62-66: every map resized bilinearly to out_hw, concatenated
    over channels. layout 'nhwc' -> [B, H*W, Ctot] (feeds `find_nn_source_correspondences`), 'nchw' -> (B,Ctot,H,W)."""
    maps = []
    for f in (feats.values() if isinstance(feats, dict) else feats):
        B, C, h, w = f.shape
        m = f.permute(0, 2, 3, 1)                      # back to the arena's token-major layout (no copy for views)
        maps.append(m.reshape(B, h * w, C).contiguous())
    r = ops.resize_concat(maps, out_hw, nhwc=(layout == "nhwc"), nchw=(layout == "nchw"))
    return r[layout]


def points_to_idxs(points, load_size):
    """correspondence_utils.py:140-146 (points in (y, x) order, numpy half-to-even rounding)."""
    points_y = np.clip(points[:, 0], 0, load_size[1] - 1)
    points_x = np.clip(points[:, 1], 0, load_size[0] - 1)
    return load_size[1] * np.round(points_y) + np.round(points_x)


def find_nn_source_correspondences(img1_feats, img2_feats, source_points, output_size, load_size):
    """Same signature / return as correspondence_utils.py:113-138. img*_feats: the NHWC stacks of `build_stack`
    ([1, h*w, C] fp16 CUDA) or (1, C, h, w) tensors (converted). Returns (points1, points2) with points2 the (y, x)
    arg-max positions on the load_size grid (int64 tensor on the GPU)."""
    def nhwc(f):
        if f.dim() == 4:
            f = f.permute(0, 2, 3, 1).reshape(f.shape[0], -1, f.shape[1])
        return f.to(torch.float16).contiguous()
    s1, s2 = nhwc(img1_feats), nhwc(img2_feats)
    assert s1.shape[0] == 1 and s1.shape == s2.shape and load_size[0] == load_size[1]
    hw = int(round(s1.shape[1] ** 0.5))
    C = s1.shape[2]
    py = np.round(np.clip(source_points[:, 0], 0, load_size[1] - 1)).astype(np.int32)
    px = np.round(np.clip(source_points[:, 1], 0, load_size[0] - 1)).astype(np.int32)
    n = len(py)
    dev = s1.device
    qyx = torch.from_numpy(np.stack([py, px], axis=-1).copy()).to(dev)
    lib = _lib.load()
    ws = torch.empty(lib.gdf_correspond_workspace_floats(n, hw, C), dtype=torch.float32, device=dev)
    idx = torch.empty(n, dtype=torch.int64, device=dev)
    check(lib.gdf_correspond(ptr(s1), ptr(s2), C, hw, load_size[0], ptr(qyx), n, ptr(idx), ptr(ws), stream_ptr()))
    points2 = torch.stack([idx // load_size[0], idx % load_size[0]], dim=-1)
    return torch.from_numpy(np.asarray(source_points)), points2


class AggregationHead:
    """Forward of the reference's trainable output processor on the feature stack:
    `AggregationNetwork.out = nn.Conv2d(dim, out_dim, 3, 1, 1, bias=False)` applied to the fp32-cast stack
    (aggregation_network.py:22,97-99). Here the fp16 NHWC stack goes straight into the tcgen05 implicit-GEMM convolution
    (fp16 operands, fp32 accumulation, fp32 output): no cast pass over the 126 MB / image stack.

        head = AggregationHead(weight)           # weight: (out_dim, dim, 3, 3) fp32, the nn.Conv2d parameter
        y = head(stack, (128, 128))              # stack: [B, H*W, dim] fp16 NHWC (build_stack) -> (B, out_dim, H, W) fp32
    """

    def __init__(self, weight, device=None):
        if weight.dim() != 4 or tuple(weight.shape[2:]) != (3, 3):
            raise ValueError("AggregationNetwork.out is a 3x3 convolution: weight must be (out_dim, dim, 3, 3)")
        dev = torch.device(device) if device is not None else weight.device
        self.out_dim, self.dim = int(weight.shape[0]), int(weight.shape[1])
        if self.dim % 64 != 0:
            raise ValueError("stack channels must be a multiple of 64 (got %d)" % self.dim)
        self.w = ops.pack_conv_weight_f16(weight.to(dev))

    def __call__(self, stack_nhwc, hw):
        B, HW, C = stack_nhwc.shape
        H, W = hw
        assert HW == H * W and C == self.dim and stack_nhwc.dtype == torch.float16 and stack_nhwc.is_contiguous()
        out = torch.empty(B * HW, self.out_dim, dtype=torch.float32, device=stack_nhwc.device)
        ep = ops.make_epilogue(out_f32=out, n_out=self.out_dim, in_f16=True)
        ops.conv3x3(stack_nhwc.view(B, H, W, C), self.w, ep)
        return out.view(B, H, W, self.out_dim).permute(0, 3, 1, 2)
