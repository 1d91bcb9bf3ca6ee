"""Forward of the segmentation consumer's trainable blocks on the captured maps:
`ResBlock` / `MultiRes` and `DiffusionSegmentor.extract_feat` of segmentation/models/diffusion_segmentor.py:23-53,
209-246, inference only (eval-mode BatchNorm2d, folded into the convolutions when the parameters are loaded).

The reference casts every fp16 map to fp32 and runs cuDNN convolutions. Here the fp16 NHWC maps of the feature arena go
straight into the tcgen05 implicit-GEMM convolution as fp16 operands (fp32 accumulation): no cast pass, no layout
change; conv1 writes relu(.) as fp16, conv2 adds the fp16 input back in its epilogue and writes either the fp16 channel
slice of the level's concatenation buffer or the fp32 result."""
import torch

from . import ops


def layer_conv_name(layer, model_index=None):
    """diffusion_segmentor.py:188-192."""
    name = layer.replace("-", "_")
    return name if model_index is None else "%d_%s" % (model_index, name)


def _fold(sd, prefix, c, eps, dev):
    w = sd["%s.%s.0.weight" % (prefix, c)].to(dev, torch.float32)
    b = sd["%s.%s.0.bias" % (prefix, c)].to(dev, torch.float32)
    g = sd["%s.%s.1.weight" % (prefix, c)].to(dev, torch.float32)
    beta = sd["%s.%s.1.bias" % (prefix, c)].to(dev, torch.float32)
    mean = sd["%s.%s.1.running_mean" % (prefix, c)].to(dev, torch.float32)
    var = sd["%s.%s.1.running_var" % (prefix, c)].to(dev, torch.float32)
    s = g * torch.rsqrt(var + eps)
    return ops.pack_conv_weight_f16(w * s[:, None, None, None]), ((b - mean) * s + beta).contiguous()


class ResBlock:
    """x + BN(conv2(relu(BN(conv1(x))))), diffusion_segmentor.py:23-44, with the BatchNorm running statistics folded.

        blk = ResBlock(state_dict, 'up_level0_upsampler_out', dim)      # the reference module's own parameter names
        y = blk(x)                        # x: [B, H, W, dim] fp16 NHWC contiguous -> fp32 [B, H, W, dim]
    """

    def __init__(self, sd, prefix, dim=None, device="cuda", eps=1e-5):
        dev = torch.device(device)
        self.dim = int(sd["%s.conv1.0.weight" % prefix].shape[0])
        if dim is not None and dim != self.dim:
            raise ValueError("ResBlock %s: %d channels in the parameters, %d expected" % (prefix, self.dim, dim))
        if self.dim % 64 != 0:
            raise ValueError("ResBlock channels must be a multiple of 64 (got %d)" % self.dim)
        self.w1, self.b1 = _fold(sd, prefix, "conv1", eps, dev)
        self.w2, self.b2 = _fold(sd, prefix, "conv2", eps, dev)

    def __call__(self, x, out_f16=None, want_f32=True):
        """out_f16: optional fp16 destination [B*H*W, >= dim] view (row pitch = stride(0): a channel slice of a
        concatenation buffer). Returns the fp32 result [B, H, W, dim] when want_f32, else out_f16."""
        B, H, W, C = x.shape
        assert C == self.dim and x.dtype == torch.float16 and x.is_contiguous()
        M = B * H * W
        h = torch.empty(M, C, dtype=torch.float16, device=x.device)
        ops.conv3x3(x, self.w1, ops.make_epilogue(out=h, bias=self.b1, act=ops.ACT_RELU, out_f16_from=-1, in_f16=True))
        y32 = torch.empty(M, C, dtype=torch.float32, device=x.device) if want_f32 else None
        ep = ops.make_epilogue(out=out_f16, out_f32=y32, bias=self.b2, residual=x.view(M, C), res_f16=True,
                               out_f16_from=-1, in_f16=True)
        ops.conv3x3(h.view(B, H, W, C), self.w2, ep)
        return y32.view(B, H, W, C) if want_f32 else out_f16


class MultiRes:
    """diffusion_segmentor.py:46-53: `nn.ModuleList([ResBlock(dim)] * n)` is ONE ResBlock applied n times."""

    def __init__(self, sd, prefix, n, dim=None, device="cuda"):
        self.block = ResBlock(sd, prefix + ".res.0", dim, device)
        self.n = n

    def __call__(self, x):
        B, H, W, C = x.shape
        for i in range(self.n):
            last = i == self.n - 1
            nxt = None if last else torch.empty(B * H * W, C, dtype=torch.float16, device=x.device)
            y = self.block(x, out_f16=nxt, want_f32=last)
            x = y if last else nxt.view(B, H, W, C)
        return x


class SegmentorFeatureHead:
    """`DiffusionSegmentor.extract_feat`, single-extractor branch (diffusion_segmentor.py:232-246), after the
    `FeatureExtractor.extract` call: per level a ResBlock per captured map, channel concat, ResBlock over the sum.

        head = SegmentorFeatureHead(feature_layers, state_dict)     # names as the reference registers them
        outs = head(features)       # features: dict id -> fp16 (B, C, h, w) views from FeatureExtractor.extract
                                    # -> list of fp32 (B, sum_dim, h, w), one per level
    """

    def __init__(self, feature_layers, sd, device="cuda"):
        self.feature_layers = feature_layers
        self.blocks, self.sums = {}, []
        for level, res in enumerate(feature_layers):
            for layer in res:
                self.blocks[layer[0]] = ResBlock(sd, layer_conv_name(layer[0]), layer[1], device)
            self.sums.append(ResBlock(sd, layer_conv_name("sum%d" % level), sum(l[1] for l in res), device))

    def __call__(self, features):
        outs = []
        for level, res in enumerate(self.feature_layers):
            f0 = features[res[0][0]]
            B, _, H, W = f0.shape
            sum_dim = self.sums[level].dim
            cat = torch.empty(B * H * W, sum_dim, dtype=torch.float16, device=f0.device)
            off = 0
            for layer in res:
                f = features[layer[0]]
                if tuple(f.shape) != (B, layer[1], H, W):
                    raise ValueError("level %d: map %s has shape %s, expected %s" % (level, layer[0], tuple(f.shape),
                                                                                      (B, layer[1], H, W)))
                x = f.permute(0, 2, 3, 1).contiguous()        # the arena's token-major layout (no copy for views)
                self.blocks[layer[0]](x, out_f16=cat[:, off:off + layer[1]], want_f32=False)
                off += layer[1]
            y = self.sums[level](cat.view(B, H, W, sum_dim))
            outs.append(y.permute(0, 3, 1, 2))
        return outs

    extract_feat = __call__
