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

    def __call__(self, x, out_f16=None, want_f32=True):
        """out_f16 / want_f32 as in ResBlock.__call__, for the LAST application."""
        B, H, W, C = x.shape
        for i in range(self.n):
            last = i == self.n - 1
            if last:
                return self.block(x, out_f16=out_f16, want_f32=want_f32)
            nxt = torch.empty(B * H * W, C, dtype=torch.float16, device=x.device)
            self.block(x, out_f16=nxt, want_f32=False)
            x = nxt.view(B, H, W, C)


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


class MultiSegmentorFeatureHead:
    """`DiffusionSegmentor.extract_feat`, several-extractors branch (diffusion_segmentor.py:248-297), after the
    `extract` calls: per model i and level a `MultiRes(dim, 4)` per captured map, channel concat, `MultiRes(sum, 2)`;
    then per level the models' results are concatenated and go through `ResBlock(c_per_level[level])` ('amalgemated').

        head = MultiSegmentorFeatureHead([layers_model0, layers_model1], c_per_level, state_dict)
        outs = head([features_model0, features_model1])        # -> list of fp32 (B, c_per_level[l], h, w)
    """

    def __init__(self, feature_layers, c_per_level, sd, device="cuda"):
        self.feature_layers = feature_layers
        self.c_per_level = list(c_per_level)
        self.blocks, self.sums = {}, {}
        for i, layers in enumerate(feature_layers):
            for level, res in enumerate(layers):
                for layer in res:
                    self.blocks[(i, layer[0])] = MultiRes(sd, layer_conv_name(layer[0], i), 4, layer[1], device)
                if res:
                    self.sums[(i, level)] = MultiRes(sd, layer_conv_name("sum%d" % level, i), 2, sum(l[1] for l in res),
                                                     device)
        self.amalgamated = [ResBlock(sd, layer_conv_name("amalgemated", l), c, device) for l, c in enumerate(c_per_level)]

    def __call__(self, features_per_model):
        outs = []
        for level, c_total in enumerate(self.c_per_level):
            parts = [(i, res[level]) for i, res in enumerate(self.feature_layers) if level < len(res) and res[level]]
            f0 = features_per_model[parts[0][0]][parts[0][1][0][0]]
            B, _, H, W = f0.shape
            cat = torch.empty(B * H * W, c_total, dtype=torch.float16, device=f0.device)
            off = 0
            for i, res in parts:
                sum_dim = sum(l[1] for l in res)
                per = torch.empty(B * H * W, sum_dim, dtype=torch.float16, device=f0.device)
                o2 = 0
                for layer in res:
                    f = features_per_model[i][layer[0]]
                    if tuple(f.shape) != (B, layer[1], H, W):
                        raise ValueError("model %d level %d: map %s has shape %s, expected %s"
                                         % (i, level, layer[0], tuple(f.shape), (B, layer[1], H, W)))
                    self.blocks[(i, layer[0])](f.permute(0, 2, 3, 1).contiguous(), out_f16=per[:, o2:o2 + layer[1]],
                                               want_f32=False)
                    o2 += layer[1]
                self.sums[(i, level)](per.view(B, H, W, sum_dim), out_f16=cat[:, off:off + sum_dim], want_f32=False)
                off += sum_dim
            if off != c_total:
                raise ValueError("level %d: the models bring %d channels, c_per_level says %d" % (level, off, c_total))
            outs.append(self.amalgamated[level](cat.view(B, H, W, c_total)).permute(0, 3, 1, 2))
        return outs

    extract_feat = __call__
