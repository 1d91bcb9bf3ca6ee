"""CPU oracle of the extraction hot path: a plain-PyTorch fp32 restatement of the reference's arithmetic.

TEST INFRASTRUCTURE ONLY. Nothing under generic_diffusion_feature_b200/ may import this file; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do. It is the checker, never the
product path.

Pinning status: the reference ships no tests or golden vectors and its arithmetic lives in
diffusers==0.32.2, which is not installed here. The modules the reference vendors (feature/diffusers/models/
resnet.py, attention.py, attention_processor.py, transformers/transformer_2d.py, downsampling.py,
upsampling.py, unet/unet_2d_condition.py) ARE executed in this container through an import shim
(tools/make_golden.py) and this restatement is checked against them module by module and end to end; the
resulting vectors are committed under tests/golden/. The un-vendored pieces (unet_2d_blocks wiring,
embeddings.Timesteps/TimestepEmbedding, activations.GEGLU, AutoencoderKL encoder and decoder, scheduler tables and
the first `step()` of Euler / PNDM) are restated from the published diffusers 0.32.2 semantics -> for those: PARITY
UNPINNED. The segmentation heads (seg_*) ARE pinned: tests/golden/segmentor_head.pt comes from the reference's own
ResBlock / MultiRes / DiffusionSegmentor.extract_feat.

Every class cites the reference file:line it follows. Attribute names mirror diffusers so that (a) the
reference's own prepare_feature_extractor (feature/components/feature_extractor.py:92-288) can attach its
FeatureGatherers to this tree unchanged and (b) state_dict keys equal the diffusers parameter names.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------ capture plumbing
class FeatureStore:
    """feature/components/feature_extractor.py:8-80 with train_unet=True semantics (no fp16 / cuda cast), so
    the oracle returns fp32 maps. Filtering, cross-k/v drop, (b (h w) c -> b c h w) and insertion order kept."""

    def __init__(self, to_store, resize_ratio=1):
        self.to_store = dict(to_store) if to_store else {}
        self.accept_all = not to_store
        self.feats = {}
        self.resize_ratio = resize_ratio

    def reset(self):
        self.feats = {}

    def store(self, feat, feat_id):
        if (feat_id in self.to_store and self.to_store[feat_id]) or self.accept_all:
            if "cross-k" in feat_id or "cross-v" in feat_id:   # :38-39
                return
            if feat.dim() == 3:                                  # :46-48
                size = int(math.sqrt(feat.shape[1]))
                feat = feat.reshape(feat.shape[0], size, size, feat.shape[2]).permute(0, 3, 1, 2)
            if self.resize_ratio > 1:                            # :51-53
                feat = F.adaptive_avg_pool2d(feat, (feat.shape[2] // self.resize_ratio, feat.shape[3] // self.resize_ratio))
            self.feats[feat_id] = feat.detach().clone()          # TF.normalize(mean=0,std=1) is an identity clone :56

    @property
    def stored_feats(self):
        return self.feats


class FeatureGatherer:
    """feature_extractor.py:83-89"""

    def __init__(self, module_id, feature_store):
        self.module_id = module_id
        self.feature_store = feature_store

    def gather(self, feat, feat_id):
        self.feature_store.store(feat, "-".join([self.module_id, feat_id]))


def attach_gatherers(unet, store):
    """Id grammar of prepare_feature_extractor for UNets (feature_extractor.py:125-249), restated."""
    unet.feature_gatherer = FeatureGatherer("unet", store)

    def do_level(prefix, level, samplers_attr, sampler_name):
        for j in range(len(level.resnets)):
            level.resnets[j].feature_gatherer = FeatureGatherer("-".join(prefix + ["repeat%d" % j, "res"]), store)
            if hasattr(level, "attentions") and len(level.attentions) > 0:
                vit = level.attentions[j]
                vit.feature_gatherer = FeatureGatherer("-".join(prefix + ["repeat%d" % j, "vit"]), store)
                for k, blk in enumerate(vit.transformer_blocks):
                    bid = prefix + ["repeat%d" % j, "vit", "block%d" % k]
                    blk.feature_gatherer = FeatureGatherer("-".join(bid), store)
                    blk.attn1.feature_gatherer = FeatureGatherer("-".join(bid + ["self"]), store)
                    blk.attn2.feature_gatherer = FeatureGatherer("-".join(bid + ["cross"]), store)
                    blk.ff.feature_gatherer = FeatureGatherer("-".join(bid + ["ffn"]), store)
        samplers = getattr(level, samplers_attr, None)
        if samplers:
            for s in samplers:
                s.feature_gatherer = FeatureGatherer("-".join(prefix + [sampler_name]), store)

    for i, level in enumerate(unet.down_blocks):
        do_level(["down", "level%d" % i], level, "downsamplers", "downsampler")
    mid = unet.mid_block
    for j in range(len(mid.resnets)):
        mid.resnets[j].feature_gatherer = FeatureGatherer("mid-repeat%d-res" % j, store)
    mid.attentions[0].feature_gatherer = FeatureGatherer("mid-vit", store)
    for k, blk in enumerate(mid.attentions[0].transformer_blocks):
        bid = ["mid", "vit", "block%d" % k]
        blk.feature_gatherer = FeatureGatherer("-".join(bid), store)
        blk.attn1.feature_gatherer = FeatureGatherer("-".join(bid + ["self"]), store)
        blk.attn2.feature_gatherer = FeatureGatherer("-".join(bid + ["cross"]), store)
        blk.ff.feature_gatherer = FeatureGatherer("-".join(bid + ["ffn"]), store)
    for i, level in enumerate(unet.up_blocks):
        do_level(["up", "level%d" % i], level, "upsamplers", "upsampler")


def _gather(mod, x, tag):
    if hasattr(mod, "feature_gatherer"):
        mod.feature_gatherer.gather(x, tag)


# ------------------------------------------------------------------------------------------ building blocks
class ResnetBlock2D(nn.Module):
    """feature/diffusers/models/resnet.py:320-379 (time_embedding_norm='default', output_scale_factor=1)."""

    def __init__(self, cin, cout, temb_ch, groups=32, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_ch, cout) if temb_ch else None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        _gather(self, h, "increment")
        out = x + h
        _gather(self, out, "out")
        return out


class Downsample2D(nn.Module):
    """feature/diffusers/models/downsampling.py:132-152; padding=0 -> F.pad (0,1,0,1) (VAE encoder)."""

    def __init__(self, ch, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1))
        x = self.conv(x)
        _gather(self, x, "out")
        return x


class Upsample2D(nn.Module):
    """feature/diffusers/models/upsampling.py:142-195 (nearest x2 then conv3x3)."""

    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        x = self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        _gather(self, x, "out")
        return x


class Attention(nn.Module):
    """feature/diffusers/models/attention_processor.py:105-297 (ctor) + AttnProcessor2_0 :3244-3331.
    q/k/v are gathered before the head split (:3291-3294)."""

    def __init__(self, dim, heads, ctx_dim=None, bias=False, out_bias=True):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=bias)
        self.to_k = nn.Linear(ctx_dim or dim, dim, bias=bias)
        self.to_v = nn.Linear(ctx_dim or dim, dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim, bias=out_bias), nn.Identity()])

    def forward(self, x, ctx=None):
        B, N, C = x.shape
        ctx = x if ctx is None else ctx
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        _gather(self, q, "q")
        _gather(self, k, "k")
        _gather(self, v, "v")
        d = C // self.heads
        q = q.view(B, -1, self.heads, d).transpose(1, 2)
        k = k.view(B, -1, self.heads, d).transpose(1, 2)
        v = v.view(B, -1, self.heads, d).transpose(1, 2)
        proc = getattr(self, "attn_store_processor", None)
        if proc is None:
            o = F.scaled_dot_product_attention(q, k, v)
        else:
            # AttnStoreProcessor (feature/components/attention.py:165-263): explicit probabilities
            # softmax(scale * Q K^T) (get_attention_scores, attention_processor.py:640-685), head-mean handed to the
            # AttentionStore (:241-242), the (B, heads, Nq, Nk) tensor gathered as `map` (:243-244), then P V (:247)
            attnstore, place = proc
            probs = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, dim=-1)
            if attnstore is not None:
                attnstore(probs.mean(1), ctx is not x, place)
            _gather(self, probs, "map")
            o = probs @ v
        o = o.transpose(1, 2).reshape(B, -1, C)
        return self.to_out[0](o)


class AttentionStore:
    """feature/components/attention.py:102-161 (only what one forward uses): head-mean maps whose query count lies in
    [min_size^2, max_size^2] are kept per "<place>_<cross|self>" category; aggregate_attention groups the maps of the
    selected categories by size, (b (h w) c -> b c h w), and averages each group."""

    KEYS = ("down_cross", "mid_cross", "up_cross", "down_self", "mid_self", "up_self")

    def __init__(self, min_size=32, max_size=64):
        self.min_size, self.max_size = min_size, max_size
        self.reset()

    def reset(self):
        self.step_store = {k: [] for k in self.KEYS}

    def __call__(self, attn, is_cross, place_in_unet):
        key = "%s_%s" % (place_in_unet, "cross" if is_cross else "self")
        if self.min_size ** 2 <= attn.shape[1] <= self.max_size ** 2:
            self.step_store[key].append(attn.detach())
        return attn

    def aggregate_attention(self, attn_selector):
        attns = {key: {} for key in attn_selector}
        for k in attn_selector:
            for a in self.step_store[k]:
                size = int(math.sqrt(a.shape[1]))
                r = a.reshape(a.shape[0], size, size, a.shape[2]).permute(0, 3, 1, 2)
                attns[k].setdefault(size, []).append(r)
            for size, a in attns[k].items():
                attns[k][size] = torch.stack(a).mean(0)
        return attns


def register_attention_store(unet, img_size, processor_only=False):
    """register_attention_store, UNet branch (feature/components/attention.py:531-566): every attention module of the
    down / mid / up blocks gets the storing processor; AttentionStore(img_size // 32, img_size // 16)."""
    store = None if processor_only else AttentionStore(img_size // 32, img_size // 16)
    for place, blocks in (("down", unet.down_blocks), ("mid", [unet.mid_block]), ("up", unet.up_blocks)):
        for blk in blocks:
            for vit in (getattr(blk, "attentions", None) or []):
                for tb in vit.transformer_blocks:
                    tb.attn1.attn_store_processor = (store, place)
                    tb.attn2.attn_store_processor = (store, place)
    return store


def aggregated_attention_feature(store, categories, img_size):
    """diffusion_feature.py:488-500: every (category, size) mean map nearest-resized to (img/8, img/8), concatenated
    on the channel axis (key index of the attention) -> the `attn` entry of the returned dict."""
    all_attns = []
    for category, maps in store.aggregate_attention(categories).items():
        for size, attn in maps.items():
            all_attns.append(F.interpolate(attn, size=(img_size // 8, img_size // 8)))
    return torch.cat(all_attns, dim=-3)


class GEGLU(nn.Module):
    """[diffusers 0.32.2 activations.GEGLU, un-vendored]: h, g = proj(x).chunk(2); h * gelu_erf(g)."""

    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, g = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(g)


class FeedForward(nn.Module):
    """feature/diffusers/models/attention.py:1209-1258; `inner` gathered after net[0] (:1253-1257)."""

    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for i, m in enumerate(self.net):
            x = m(x)
            if i == 0:
                _gather(self, x, "inner")
        return x


class BasicTransformerBlock(nn.Module):
    """feature/diffusers/models/attention.py:469-592, norm_type='layer_norm'."""

    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), ctx) + x
        x = self.ff(self.norm3(x)) + x
        _gather(self, x, "out")
        return x


class Transformer2DModel(nn.Module):
    """feature/diffusers/models/transformers/transformer_2d.py:174-209 (ctor), :403-475 (forward),
    :482-530 (continuous in/out). GroupNorm eps 1e-6."""

    def __init__(self, dim, heads, depth, ctx_dim, linear_proj, groups=32):
        super().__init__()
        self.use_linear_projection = linear_proj
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, ctx_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim) if linear_proj else nn.Conv2d(dim, dim, 1)

    def forward(self, x, ctx):
        B, C, H, W = x.shape
        res = x
        h = self.norm(x)
        if self.use_linear_projection:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(B, H * W, C))
        else:
            h = self.proj_in(h).permute(0, 2, 3, 1).reshape(B, H * W, C)
        for blk in self.transformer_blocks:
            h = blk(h, ctx)
        if self.use_linear_projection:
            h = self.proj_out(h).reshape(B, H, W, C).permute(0, 3, 1, 2)
        else:
            h = self.proj_out(h.reshape(B, H, W, C).permute(0, 3, 1, 2))
        out = h + res
        _gather(self, out, "out")
        return out


class DownBlock(nn.Module):
    """[diffusers unet_2d_blocks.CrossAttnDownBlock2D / DownBlock2D, un-vendored] SURVEY Appendix C."""

    def __init__(self, cin, cout, temb, n_layers, has_attn, heads, depth, ctx_dim, linear_proj, add_down, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, eps=eps)
                                      for i in range(n_layers)])
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, depth, ctx_dim, linear_proj)
                                             for _ in range(n_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, h, temb, ctx):
        skips = []
        for i, r in enumerate(self.resnets):
            h = r(h, temb)
            if hasattr(self, "attentions"):
                h = self.attentions[i](h, ctx)
            skips.append(h)
        if self.downsamplers:
            h = self.downsamplers[0](h)
            skips.append(h)
        return h, skips


class MidBlock(nn.Module):
    """[UNetMidBlock2DCrossAttn, un-vendored]: resnets[0] -> (attn -> resnet)*."""

    def __init__(self, ch, temb, heads, depth, ctx_dim, linear_proj, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb, eps=eps), ResnetBlock2D(ch, ch, temb, eps=eps)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, heads, depth, ctx_dim, linear_proj)])

    def forward(self, h, temb, ctx):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, ctx)
        return self.resnets[1](h, temb)


class UpBlock(nn.Module):
    """[CrossAttnUpBlock2D / UpBlock2D, un-vendored]: h = cat([h, skips.pop()], 1); resnet; attn; upsampler."""

    def __init__(self, cin_prev, cout, skip_chs, temb, has_attn, heads, depth, ctx_dim, linear_proj, add_up, eps):
        super().__init__()
        self.resnets = nn.ModuleList()
        for i, sc in enumerate(skip_chs):
            self.resnets.append(ResnetBlock2D((cin_prev if i == 0 else cout) + sc, cout, temb, eps=eps))
        if has_attn:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, depth, ctx_dim, linear_proj)
                                             for _ in skip_chs])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, h, skips, temb, ctx):
        for i, r in enumerate(self.resnets):
            h = torch.cat([h, skips.pop()], dim=1)
            h = r(h, temb)
            if hasattr(self, "attentions"):
                h = self.attentions[i](h, ctx)
        if self.upsamplers:
            h = self.upsamplers[0](h)
        return h


def timestep_embedding(t, dim):
    """[diffusers embeddings.get_timestep_embedding, flip_sin_to_cos=True, downscale_freq_shift=0]."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    e = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(e), torch.sin(e)], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


UNET_CONFIGS = {
    # SURVEY.md Appendix B
    "xl": dict(block_out=(320, 640, 1280), down_attn=(0, 1, 1), up_attn=(1, 1, 0), depth=(1, 2, 10),
               heads=(5, 10, 20), ctx_dim=2048, linear_proj=True, add_time_dim=256, add_in=2816, eps=1e-5),
    "1-5": dict(block_out=(320, 640, 1280, 1280), down_attn=(1, 1, 1, 0), up_attn=(0, 1, 1, 1), depth=(1, 1, 1, 1),
                heads=(8, 8, 8, 8), ctx_dim=768, linear_proj=False, add_time_dim=0, add_in=0, eps=1e-5),
    "2-1": dict(block_out=(320, 640, 1280, 1280), down_attn=(1, 1, 1, 0), up_attn=(0, 1, 1, 1), depth=(1, 1, 1, 1),
                heads=(5, 10, 20, 20), ctx_dim=1024, linear_proj=True, add_time_dim=0, add_in=0, eps=1e-5),
}


class UNet2DConditionModel(nn.Module):
    """feature/diffusers/models/unet/unet_2d_condition.py:171-484 (ctor), :1040-1319 (forward)."""

    def __init__(self, cfg, layers_per_block=2):
        super().__init__()
        self.cfg = cfg
        bo = cfg["block_out"]
        temb = bo[0] * 4
        self.conv_in = nn.Conv2d(4, bo[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(bo[0], temb)
        if cfg["add_time_dim"]:
            self.add_embedding = TimestepEmbedding(cfg["add_in"], temb)
        self.down_blocks = nn.ModuleList()
        n = len(bo)
        ch = bo[0]
        skip_chs = [bo[0]]
        for i in range(n):
            cin, ch = ch, bo[i]
            last = i == n - 1
            self.down_blocks.append(DownBlock(cin, ch, temb, layers_per_block, cfg["down_attn"][i], cfg["heads"][i],
                                              cfg["depth"][i], cfg["ctx_dim"], cfg["linear_proj"], not last,
                                              cfg["eps"]))
            skip_chs += [ch] * layers_per_block + ([ch] if not last else [])
        self.mid_block = MidBlock(bo[-1], temb, cfg["heads"][-1], cfg["depth"][-1], cfg["ctx_dim"],
                                  cfg["linear_proj"], cfg["eps"])
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(bo))
        rheads, rdepth = list(reversed(cfg["heads"])), list(reversed(cfg["depth"]))
        prev = bo[-1]
        for i in range(n):
            cout = rev[i]
            sk = [skip_chs.pop() for _ in range(layers_per_block + 1)]
            self.up_blocks.append(UpBlock(prev, cout, sk, temb, cfg["up_attn"][i], rheads[i], rdepth[i],
                                          cfg["ctx_dim"], cfg["linear_proj"], i != n - 1, cfg["eps"]))
            prev = cout
        self.conv_norm_out = nn.GroupNorm(32, bo[0], eps=cfg["eps"])
        self.conv_out = nn.Conv2d(bo[0], 4, 3, padding=1)

    def forward(self, sample, timestep, ctx, text_embeds=None, time_ids=None, down_residuals=None, mid_residual=None):
        B = sample.shape[0]
        t = torch.as_tensor(timestep, dtype=torch.float32).reshape(-1).expand(B)
        emb = self.time_embedding(timestep_embedding(t, self.cfg["block_out"][0]))        # :1141-1142
        if self.cfg["add_time_dim"]:                                                      # :968-984 (text_time)
            te = timestep_embedding(time_ids.flatten(), self.cfg["add_time_dim"]).reshape(B, -1)
            emb = emb + self.add_embedding(torch.cat([text_embeds, te], dim=-1))
        _gather(self, sample, "in")                                                       # :1169-1173
        h = self.conv_in(sample)
        _gather(self, h, "after-conv-in")
        skips = [h]
        for blk in self.down_blocks:
            h, s = blk(h, emb, ctx)
            skips += s
        if down_residuals is not None:        # ControlNet: one residual per skip tensor (:1244-1252) ...
            skips = [s + r for s, r in zip(skips, down_residuals)]
        h = self.mid_block(h, emb, ctx)
        if mid_residual is not None:          # ... and one after the mid block (:1274-1275)
            h = h + mid_residual
        for blk in self.up_blocks:
            h = blk(h, skips, emb, ctx)
        h = self.conv_out(F.silu(self.conv_norm_out(h)))                                  # :1304-1307
        _gather(self, h, "out")
        return h


# ------------------------------------------------------------------------------------------ PixArt DiT
DIT_CONFIGS = {
    # [PixArt-alpha/PixArt-Sigma-XL-2-{1024,512}-MS transformer/config.json, from memory; SURVEY.md row a16]
    "pixart-sigma": dict(layers=28, heads=16, head_dim=72, in_ch=4, out_ch=8, patch=2, caption_dim=4096,
                         sample_size=128, interpolation_scale=2.0, eps=1e-6),
    "pixart-sigma-512": dict(layers=28, heads=16, head_dim=72, in_ch=4, out_ch=8, patch=2, caption_dim=4096,
                             sample_size=64, interpolation_scale=1.0, eps=1e-6),
}


def sincos_pos_embed_2d(dim, grid, base_size, interpolation_scale):
    """[diffusers 0.32.2 embeddings.get_2d_sincos_pos_embed, un-vendored]: grid positions
    arange(grid) / (grid / base_size) / interpolation_scale; first half of the channels encodes the x (width)
    coordinate, second half y (np.meshgrid(grid_w, grid_h) puts w first); each half = [sin | cos] over
    omega_k = 10000^(-k / (dim/4)). float64 like numpy, returned fp32 (grid*grid, dim)."""
    pos = torch.arange(grid, dtype=torch.float32).double() / (grid / base_size) / interpolation_scale
    gw = pos[None, :].expand(grid, grid).reshape(-1)      # x varies fastest
    gh = pos[:, None].expand(grid, grid).reshape(-1)
    quarter = dim // 4
    omega = 1.0 / 10000 ** (torch.arange(quarter, dtype=torch.float64) / quarter)

    def one(p):
        o = p[:, None] * omega[None]
        return torch.cat([torch.sin(o), torch.cos(o)], dim=1)
    return torch.cat([one(gw), one(gh)], dim=1).float()


class PatchEmbed(nn.Module):
    """[diffusers embeddings.PatchEmbed, un-vendored; equivalent use at transformer_2d.py:541-569]:
    Conv2d(k = s = patch) -> flatten -> + fixed sin-cos table (a persistent buffer named pos_embed)."""

    def __init__(self, cfg):
        super().__init__()
        dim = cfg["heads"] * cfg["head_dim"]
        grid = cfg["sample_size"] // cfg["patch"]
        self.proj = nn.Conv2d(cfg["in_ch"], dim, cfg["patch"], stride=cfg["patch"])
        self.register_buffer("pos_embed", sincos_pos_embed_2d(dim, grid, grid, cfg["interpolation_scale"])[None])

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2) + self.pos_embed


class _TimestepEmb(nn.Module):
    """[PixArtAlphaCombinedTimestepSizeEmbeddings without additional conditions]: Timesteps(256, flip, shift 0)
    -> TimestepEmbedding(256, dim)."""

    def __init__(self, dim):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, dim)

    def forward(self, t):
        return self.timestep_embedder(timestep_embedding(t, 256))


class AdaLayerNormSingle(nn.Module):
    """[diffusers normalization.AdaLayerNormSingle, un-vendored]: (Linear(SiLU(emb)) -> 6*dim, emb)."""

    def __init__(self, dim):
        super().__init__()
        self.emb = _TimestepEmb(dim)
        self.linear = nn.Linear(dim, 6 * dim)

    def forward(self, t):
        e = self.emb(t)
        return self.linear(F.silu(e)), e


class CaptionProjection(nn.Module):
    """[diffusers embeddings.PixArtAlphaTextProjection, act gelu_tanh]."""

    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, c):
        return self.linear_2(F.gelu(self.linear_1(c), approximate="tanh"))


class MaskedAttention(Attention):
    """Attention (attention_processor.py:3244-3331) with bias on q/k/v and an additive key mask
    (attention_mask prepared as (1 - mask) * -10000 by the PixArt transformer)."""

    def forward(self, x, ctx=None, key_bias=None):
        B, N, C = x.shape
        ctx = x if ctx is None else ctx
        q, k, v = self.to_q(x), self.to_k(ctx), self.to_v(ctx)
        _gather(self, q, "q")
        _gather(self, k, "k")
        _gather(self, v, "v")
        d = C // self.heads
        q = q.view(B, -1, self.heads, d).transpose(1, 2)
        k = k.view(B, -1, self.heads, d).transpose(1, 2)
        v = v.view(B, -1, self.heads, d).transpose(1, 2)
        m = None if key_bias is None else key_bias[:, None, None, :]
        proc = getattr(self, "attn_store_processor", None)
        if proc is None:
            o = F.scaled_dot_product_attention(q, k, v, attn_mask=m)
        else:
            # AttnStoreProcessor on a PixArt block (feature/components/attention.py:165-263, registered by :583-593):
            # get_attention_scores = softmax(mask + scale * Q K^T) (baddbmm with the prepared additive mask)
            attnstore, place = proc
            sc = q @ k.transpose(-1, -2) * d ** -0.5
            probs = torch.softmax(sc if m is None else sc + m, dim=-1)
            if attnstore is not None:
                attnstore(probs.mean(1), ctx is not x, place)
            _gather(self, probs, "map")
            o = probs @ v
        o = o.transpose(1, 2).reshape(B, -1, C)
        return self.to_out[0](o)


class GELUProj(nn.Module):
    """[diffusers activations.GELU(approximate='tanh'), un-vendored]."""

    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner)

    def forward(self, x):
        return F.gelu(self.proj(x), approximate="tanh")


class FeedForwardGelu(nn.Module):
    """attention.py:1209-1258 with activation_fn='gelu-approximate'; `inner` gathered after net[0]."""

    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GELUProj(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for i, m in enumerate(self.net):
            x = m(x)
            if i == 0:
                _gather(self, x, "inner")
        return x


class AdaSingleTransformerBlock(nn.Module):
    """feature/diffusers/models/attention.py:469-592 with norm_type='ada_norm_single' (:498-503 modulation,
    :523-524 gate, :539-542 no norm before cross-attention, :570-583 norm2 + mlp modulation + gate)."""

    def __init__(self, dim, heads, eps):
        super().__init__()
        self.scale_shift_table = nn.Parameter(torch.zeros(6, dim))
        self.norm1 = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.attn1 = MaskedAttention(dim, heads, bias=True)
        self.norm2 = nn.LayerNorm(dim, eps=eps, elementwise_affine=False)
        self.attn2 = MaskedAttention(dim, heads, dim, bias=True)
        self.ff = FeedForwardGelu(dim)

    def forward(self, x, ctx, key_bias, t6):
        B = x.shape[0]
        sh_a, sc_a, g_a, sh_m, sc_m, g_m = (self.scale_shift_table[None] + t6.reshape(B, 6, -1)).chunk(6, dim=1)
        x = g_a * self.attn1(self.norm1(x) * (1 + sc_a) + sh_a) + x
        x = self.attn2(x, ctx, key_bias) + x
        x = g_m * self.ff(self.norm2(x) * (1 + sc_m) + sh_m) + x
        _gather(self, x, "out")
        return x


class PixArtTransformer2DModel(nn.Module):
    """[diffusers 0.32.2 models/transformers/pixart_transformer_2d.py, un-vendored; the equivalent patch path is
    visible in the reference at transformers/transformer_2d.py:497-515,541-569]. Parameter names equal diffusers'.
    Called by the reference at diffusion_feature.py:467-474 with added_cond_kwargs resolution/aspect_ratio None."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        dim = cfg["heads"] * cfg["head_dim"]
        self.pos_embed = PatchEmbed(cfg)
        self.adaln_single = AdaLayerNormSingle(dim)
        self.caption_projection = CaptionProjection(cfg["caption_dim"], dim)
        self.transformer_blocks = nn.ModuleList([AdaSingleTransformerBlock(dim, cfg["heads"], cfg["eps"])
                                                 for _ in range(cfg["layers"])])
        self.norm_out = nn.LayerNorm(dim, eps=cfg["eps"], elementwise_affine=False)
        self.scale_shift_table = nn.Parameter(torch.zeros(2, dim))
        self.proj_out = nn.Linear(dim, cfg["patch"] * cfg["patch"] * cfg["out_ch"])

    def forward(self, sample, timestep, ctx, ctx_mask=None):
        B, _, H, W = sample.shape
        p, oc = self.cfg["patch"], self.cfg["out_ch"]
        key_bias = None if ctx_mask is None else (1 - ctx_mask.to(sample.dtype)) * -10000.0
        x = self.pos_embed(sample)
        t = torch.as_tensor(timestep, dtype=torch.float32).reshape(-1).expand(B)
        t6, emb = self.adaln_single(t)
        c = self.caption_projection(ctx)
        for blk in self.transformer_blocks:
            x = blk(x, c, key_bias, t6)
        shift, scale = (self.scale_shift_table[None] + emb[:, None]).chunk(2, dim=1)
        x = self.proj_out(self.norm_out(x) * (1 + scale) + shift)
        h, w = H // p, W // p
        x = x.reshape(B, h, w, p, p, oc)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(B, oc, h * p, w * p)


def register_attention_store_dit(model, img_size, processor_only=False):
    """register_attention_store, transformer branch (feature/components/attention.py:567-593): attn1 and attn2 of every
    block get the storing processor with place 'up'; AttentionStore(img_size // 32, img_size // 8)."""
    store = None if processor_only else AttentionStore(img_size // 32, img_size // 8)
    for blk in model.transformer_blocks:
        blk.attn1.attn_store_processor = (store, "up")
        blk.attn2.attn_store_processor = (store, "up")
    return store


def register_attention_store_flux(model, img_size, processor_only=False):
    """register_attention_store, Flux branch (feature/components/attention.py:567-603): the attention of every double
    and single block gets FluxAttnStoreProcessor with place 'up'; AttentionStore(img_size // 32, img_size // 8)."""
    store = None if processor_only else AttentionStore(img_size // 32, img_size // 8)
    for blk in list(model.transformer_blocks) + list(model.single_transformer_blocks):
        blk.attn.attn_store_processor = (store, "up")
    return store


def attach_gatherers_dit(model, store):
    """prepare_feature_extractor, `hasattr(pipe, 'transformer')` branch (feature_extractor.py:259-286)."""
    for i, blk in enumerate(model.transformer_blocks):
        bid = "vit-block%d" % i
        blk.feature_gatherer = FeatureGatherer(bid, store)
        blk.attn1.feature_gatherer = FeatureGatherer(bid + "-self", store)
        blk.attn2.feature_gatherer = FeatureGatherer(bid + "-cross", store)
        blk.ff.feature_gatherer = FeatureGatherer(bid + "-ffn", store)


# ------------------------------------------------------------------------------------------ Flux MMDiT
class RMSNorm(nn.Module):
    """[diffusers normalization.RMSNorm, un-vendored; built by Attention(qk_norm='rms_norm'),
    attention_processor.py:205-207, 278-280]: x * rsqrt(mean(x^2) + eps) * weight."""

    def __init__(self, dim, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x):
        v = x.float().pow(2).mean(-1, keepdim=True)
        return x * torch.rsqrt(v + self.eps) * self.weight


def flux_rope(ids, axes_dim, theta=10000.0):
    """[diffusers embeddings.FluxPosEmbed + get_1d_rotary_pos_embed(use_real=True, repeat_interleave_real=True,
    freqs_dtype=float64), un-vendored; called at transformer_flux.py:481-482]: (cos, sin), each [S, sum(axes_dim)]."""
    cos, sin = [], []
    pos = ids.double()
    for i, d in enumerate(axes_dim):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64)[: d // 2] / d))
        ang = torch.outer(pos[:, i], freqs)
        cos.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos, dim=-1), torch.cat(sin, dim=-1)


def apply_rotary_emb(x, rope):
    """[diffusers embeddings.apply_rotary_emb, use_real=True, use_real_unbind_dim=-1, un-vendored; called at
    attention_processor.py:2330-2334]: x [B, H, S, D]."""
    cos, sin = rope
    cos, sin = cos[None, None], sin[None, None]
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos + rot.float() * sin).to(x.dtype)


class FluxAttention(nn.Module):
    """Attention ctor (attention_processor.py:105-297) as built by the Flux blocks (bias, qk_norm='rms_norm' eps 1e-6;
    double blocks: added_kv_proj_dim = dim, context_pre_only False; single blocks: pre_only) + FluxAttnProcessor2_0
    (:2259-2362) including its gather call sites (:2280-2289 q / k / v, :2355-2360 attn-out)."""

    def __init__(self, dim, heads, joint):
        super().__init__()
        self.heads = heads
        d = dim // heads
        self.to_q, self.to_k, self.to_v = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
        self.norm_q, self.norm_k = RMSNorm(d, 1e-6), RMSNorm(d, 1e-6)
        self.joint = joint
        if joint:
            self.add_q_proj, self.add_k_proj, self.add_v_proj = nn.Linear(dim, dim), nn.Linear(dim, dim), nn.Linear(dim, dim)
            self.norm_added_q, self.norm_added_k = RMSNorm(d, 1e-6), RMSNorm(d, 1e-6)
            self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Identity()])
            self.to_add_out = nn.Linear(dim, dim)

    def forward(self, x, ctx=None, rope=None):
        B = x.shape[0]
        q, k, v = self.to_q(x), self.to_k(x), self.to_v(x)
        if ctx is not None:
            _gather(self, q, "q")
            _gather(self, k, "k")
            _gather(self, v, "v")
        else:
            tl = self.text_len
            _gather(self, q[:, tl:, :], "q")
            _gather(self, k[:, tl:, :], "k")
            _gather(self, v[:, tl:, :], "v")
        d = k.shape[-1] // self.heads
        sp = lambda t: t.view(B, -1, self.heads, d).transpose(1, 2)
        q, k, v = self.norm_q(sp(q)), self.norm_k(sp(k)), sp(v)
        if ctx is not None:
            cq = self.norm_added_q(sp(self.add_q_proj(ctx)))
            ck = self.norm_added_k(sp(self.add_k_proj(ctx)))
            cv = sp(self.add_v_proj(ctx))
            q, k, v = torch.cat([cq, q], dim=2), torch.cat([ck, k], dim=2), torch.cat([cv, v], dim=2)
        if rope is not None:
            q, k = apply_rotary_emb(q, rope), apply_rotary_emb(k, rope)
        proc = getattr(self, "attn_store_processor", None)
        if proc is None:
            o = F.scaled_dot_product_attention(q, k, v)
        else:
            # FluxAttnStoreProcessor (feature/components/attention.py:402-527): explicit softmax of the JOINT sequence
            # (my_scaled_dot_product_attention, :264-292); the image-query rows are split into the text keys
            # (`cross-map`) and the image keys (`self-map`), head means go to the AttentionStore (cross first), :494-502
            attnstore, place = proc
            tl = ctx.shape[1] if ctx is not None else self.text_len
            probs = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5, dim=-1)
            cross, self_ = probs[:, :, tl:, :tl], probs[:, :, tl:, tl:]
            if attnstore is not None:
                attnstore(cross.mean(1), True, place)
                attnstore(self_.mean(1), False, place)
            _gather(self, cross, "cross-map")
            _gather(self, self_, "self-map")
            o = probs @ v
        o = o.transpose(1, 2).reshape(B, -1, self.heads * d)
        if ctx is not None:
            co, o = o[:, : ctx.shape[1]], o[:, ctx.shape[1]:]
            o = self.to_out[1](self.to_out[0](o))
            co = self.to_add_out(co)
            _gather(self, o, "attn-out")
            return o, co
        _gather(self, o[:, self.text_len:, :], "attn-out")
        return o


class _AdaLNLinear(nn.Module):
    """Shared shape of [diffusers normalization.AdaLayerNormZero / AdaLayerNormZeroSingle / AdaLayerNormContinuous,
    un-vendored]: emb = Linear(SiLU(conditioning)) -> n chunks; LayerNorm without affine, eps 1e-6."""

    def __init__(self, dim, n):
        super().__init__()
        self.linear = nn.Linear(dim, n * dim)
        self.n = n
        self.dim = dim

    def chunks(self, emb):
        return self.linear(F.silu(emb)).chunk(self.n, dim=1)

    def ln(self, x):
        return F.layer_norm(x, (self.dim,), eps=1e-6)


class FluxTransformerBlock(nn.Module):
    """transformer_flux.py:116-226 (double-stream block). AdaLayerNormZero: (shift_msa, scale_msa, gate_msa, shift_mlp,
    scale_mlp, gate_mlp) = chunk(6); x = LN(x) * (1 + scale_msa) + shift_msa."""

    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = _AdaLNLinear(dim, 6)
        self.norm1_context = _AdaLNLinear(dim, 6)
        self.attn = FluxAttention(dim, heads, joint=True)
        self.ff = FeedForwardGelu(dim)
        self.ff_context = FeedForwardGelu(dim)
        self.dim = dim

    def forward(self, x, c, temb, rope):
        sh, sc, g_msa, sh_mlp, sc_mlp, g_mlp = self.norm1.chunks(temb)
        csh, csc, cg_msa, csh_mlp, csc_mlp, cg_mlp = self.norm1_context.chunks(temb)
        nx = self.norm1.ln(x) * (1 + sc[:, None]) + sh[:, None]
        nc = self.norm1_context.ln(c) * (1 + csc[:, None]) + csh[:, None]
        ao, co = self.attn(nx, nc, rope)
        x = x + g_msa.unsqueeze(1) * ao                                             # :192-193
        nx = self.norm1.ln(x) * (1 + sc_mlp[:, None]) + sh_mlp[:, None]             # :195-196 (norm2, no affine)
        _gather(self, nx, "norm-out")                                               # :200-201
        x = x + g_mlp.unsqueeze(1) * self.ff(nx)                                    # :203-206
        _gather(self, nx, "out")                                                    # :210-211 stores norm_hidden_states
        c = c + cg_msa.unsqueeze(1) * co                                            # :215-216
        ncc = self.norm1_context.ln(c) * (1 + csc_mlp[:, None]) + csh_mlp[:, None]  # :218-219
        c = c + cg_mlp.unsqueeze(1) * self.ff_context(ncc)                          # :221-222
        return c, x


class FluxSingleTransformerBlock(nn.Module):
    """transformer_flux.py:46-112. AdaLayerNormZeroSingle: (shift, scale, gate) = chunk(3)."""

    def __init__(self, dim, heads):
        super().__init__()
        self.norm = _AdaLNLinear(dim, 3)
        self.proj_mlp = nn.Linear(dim, 4 * dim)
        self.proj_out = nn.Linear(5 * dim, dim)
        self.attn = FluxAttention(dim, heads, joint=False)

    def forward(self, x, temb, rope):
        sh, sc, gate = self.norm.chunks(temb)
        nx = self.norm.ln(x) * (1 + sc[:, None]) + sh[:, None]
        mlp = F.gelu(self.proj_mlp(nx), approximate="tanh")
        ao = self.attn(nx, None, rope)
        x = x + gate.unsqueeze(1) * self.proj_out(torch.cat([ao, mlp], dim=2))
        _gather(self, x[:, self.attn.text_len:, :], "out")                          # :107-108
        return x


class _TextProj(nn.Module):
    """[diffusers embeddings.PixArtAlphaTextProjection(act_fn='silu'), un-vendored]."""

    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class _FluxTimeTextEmbed(nn.Module):
    """[diffusers embeddings.CombinedTimestepGuidanceTextProjEmbeddings / CombinedTimestepTextProjEmbeddings,
    un-vendored; built at transformer_flux.py:284-289]: time + (guidance) + pooled-text embeddings, summed."""

    def __init__(self, dim, pooled_dim, guidance):
        super().__init__()
        self.timestep_embedder = TimestepEmbedding(256, dim)
        if guidance:
            self.guidance_embedder = TimestepEmbedding(256, dim)
        self.text_embedder = _TextProj(pooled_dim, dim)
        self.guidance = guidance

    def forward(self, t, g, pooled):
        e = self.timestep_embedder(timestep_embedding(t, 256))
        if self.guidance:
            e = e + self.guidance_embedder(timestep_embedding(g, 256))
        return e + self.text_embedder(pooled)


class FluxTransformer2DModel(nn.Module):
    """transformer_flux.py:253-325 (ctor) and :414-604 (forward), parameter names equal diffusers'."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        dim = cfg["heads"] * cfg["head_dim"]
        self.time_text_embed = _FluxTimeTextEmbed(dim, cfg["pooled_dim"], cfg["guidance_embeds"])
        self.context_embedder = nn.Linear(cfg["joint_dim"], dim)
        self.x_embedder = nn.Linear(cfg["in_ch"], dim)
        self.transformer_blocks = nn.ModuleList([FluxTransformerBlock(dim, cfg["heads"]) for _ in range(cfg["layers"])])
        self.single_transformer_blocks = nn.ModuleList([FluxSingleTransformerBlock(dim, cfg["heads"])
                                                        for _ in range(cfg["single_layers"])])
        self.norm_out = _AdaLNLinear(dim, 2)
        self.proj_out = nn.Linear(dim, cfg["in_ch"])

    def forward(self, hidden_states, ctx, pooled, timestep, img_ids, txt_ids, guidance=None):
        B = hidden_states.shape[0]
        x = self.x_embedder(hidden_states)
        t = torch.as_tensor(timestep, dtype=torch.float32).reshape(-1).expand(B) * 1000          # :455
        g = None if guidance is None else torch.as_tensor(guidance, dtype=torch.float32).reshape(-1).expand(B) * 1000
        temb = self.time_text_embed(t, g, pooled)
        c = self.context_embedder(ctx)
        rope = flux_rope(torch.cat((txt_ids, img_ids), dim=0), self.cfg["axes_dims_rope"])
        for blk in self.transformer_blocks:
            c, x = blk(x, c, temb, rope)
        x = torch.cat([c, x], dim=1)                                                            # :539
        tl = c.shape[1]
        for blk in self.single_transformer_blocks:
            blk.attn.text_len = tl                                                              # :544-545
            x = blk(x, temb, rope)
        x = x[:, tl:]
        scale, shift = self.norm_out.chunks(temb)     # AdaLayerNormContinuous: scale, shift = chunk(emb, 2)
        x = self.norm_out.ln(x) * (1 + scale[:, None]) + shift[:, None]
        return self.proj_out(x)


def attach_gatherers_flux(model, store):
    """prepare_feature_extractor, `version == 'flux'` branch (feature_extractor.py:98-123)."""
    i = -1
    for i, blk in enumerate(model.transformer_blocks):
        bid = "vit-block%d" % i
        blk.feature_gatherer = FeatureGatherer(bid, store)
        blk.attn.feature_gatherer = FeatureGatherer(bid, store)
        blk.ff.feature_gatherer = FeatureGatherer(bid + "-ffn", store)
    for blk in model.single_transformer_blocks:
        i += 1
        bid = "vit-block%d" % i
        blk.feature_gatherer = FeatureGatherer(bid, store)
        blk.attn.feature_gatherer = FeatureGatherer(bid, store)


def flux_latent_image_ids(h, w):
    """pipeline_flux_img2img.py:483-494."""
    ids = torch.zeros(h, w, 3)
    ids[..., 1] = ids[..., 1] + torch.arange(h)[:, None]
    ids[..., 2] = ids[..., 2] + torch.arange(w)[None, :]
    return ids.reshape(h * w, 3)


def flux_pack_latents(lat):
    """pipeline_flux_img2img.py:498-503."""
    B, C, H, W = lat.shape
    lat = lat.view(B, C, H // 2, 2, W // 2, 2).permute(0, 2, 4, 1, 3, 5)
    return lat.reshape(B, (H // 2) * (W // 2), C * 4)


def resolve_flux_sigma(t, img_size, steps=28, base_shift=0.5, max_shift=1.15, base_seq=256, max_seq=4096):
    """pipeline_flux_img2img.py:745-766 + get_timesteps :416-425 + [FlowMatchEulerDiscreteScheduler.set_timesteps with
    use_dynamic_shifting, un-vendored]: the sigma of the first step the reference's call runs (strength = t / 1000,
    28 steps). PARITY UNPINNED (un-vendored scheduler / FLUX.1-dev scheduler_config.json from memory)."""
    import numpy as np
    strength = t / 1000
    t_start = int(max(steps - min(steps * strength, steps), 0))
    s = float(np.linspace(1.0, 1 / steps, steps)[t_start])
    seq = (img_size // 8 // 2) ** 2
    m = (max_shift - base_shift) / (max_seq - base_seq)
    mu = seq * m + (base_shift - m * base_seq)
    return float(np.float32(math.exp(mu) / (math.exp(mu) + (1 / s - 1))))


@torch.no_grad()
def extract_flux(model, vae, store, image, ctx, pooled, eps_vae, eps_q, t=50, guidance=1.0):
    """The reference's Flux path (diffusion_feature.py:246-253 -> FluxImg2ImgPipeline.__call__, pipeline_flux_img2img.py
    :700-841, returning after the first transformer call): VAE encode, (z - shift) * scale (:411), scale_noise with
    the resolved sigma (:565), pack (:566), one transformer forward with timestep = sigma (:812-822)."""
    store.reset()
    B, _, H, W = image.shape
    sigma = resolve_flux_sigma(t, H)
    m = vae.moments(image.float())
    mean, logvar = m.chunk(2, dim=1)
    logvar = logvar.clamp(-30.0, 20.0)
    z = (mean + torch.exp(0.5 * logvar) * eps_vae - vae.shift_factor) * vae.scaling_factor
    latents = sigma * eps_q + (1.0 - sigma) * z                     # FlowMatchEulerDiscreteScheduler.scale_noise
    h, w = H // 16, W // 16
    packed = flux_pack_latents(latents)
    ctx_b = ctx.expand(B, -1, -1) if ctx.shape[0] == 1 else ctx
    pooled_b = pooled.expand(B, -1) if pooled.shape[0] == 1 else pooled
    g = guidance if model.cfg["guidance_embeds"] else None
    out = model(packed, ctx_b, pooled_b, sigma, flux_latent_image_ids(h, w), torch.zeros(ctx.shape[1], 3), g)
    return store.stored_feats, latents, out


# ------------------------------------------------------------------------------------------ VAE encoder
class VaeAttention(nn.Module):
    """[diffusers Attention(512, heads=1, dim_head=512, norm_num_groups=32, residual_connection=True, bias=True)]
    through the 4-D branch of AttnProcessor2_0 (attention_processor.py:3264-3266, 3278-3279, 3323-3329)."""

    def __init__(self, ch, groups=32, eps=1e-6):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=eps)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch), nn.Identity()])

    def forward(self, x):
        B, C, H, W = x.shape
        h = self.group_norm(x).view(B, C, H * W).transpose(1, 2)
        q, k, v = self.to_q(h), self.to_k(h), self.to_v(h)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.to_out[0](o)
        return o.transpose(1, 2).reshape(B, C, H, W) + x


class VaeDownBlock(nn.Module):
    def __init__(self, cin, cout, n_layers, add_down, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, 0, eps=eps)
                                      for i in range(n_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(cout, padding=0)]) if add_down else None

    def forward(self, h):
        for r in self.resnets:
            h = r(h)
        if self.downsamplers:
            h = self.downsamplers[0](h)
        return h


class VaeMidBlock(nn.Module):
    def __init__(self, ch, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, 0, eps=eps), ResnetBlock2D(ch, ch, 0, eps=eps)])
        self.attentions = nn.ModuleList([VaeAttention(ch, eps=eps)])

    def forward(self, h):
        return self.resnets[1](self.attentions[0](self.resnets[0](h)))


class VaeEncoder(nn.Module):
    """[diffusers autoencoders/vae.Encoder + AutoencoderKL.quant_conv, un-vendored] SURVEY 8(a4)."""

    def __init__(self, block_out=(128, 256, 512, 512), layers=2, latent=4, eps=1e-6):
        super().__init__()
        self.conv_in = nn.Conv2d(3, block_out[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        ch = block_out[0]
        for i, co in enumerate(block_out):
            self.down_blocks.append(VaeDownBlock(ch, co, layers, i != len(block_out) - 1, eps))
            ch = co
        self.mid_block = VaeMidBlock(ch, eps)
        self.conv_norm_out = nn.GroupNorm(32, ch, eps=eps)
        self.conv_out = nn.Conv2d(ch, 2 * latent, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for b in self.down_blocks:
            h = b(h)
        h = self.mid_block(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class VaeUpBlock(nn.Module):
    """[diffusers unet_2d_blocks.UpDecoderBlock2D, un-vendored]: layers + 1 resnets (no time embedding), then
    Upsample2D (nearest x2 + conv3x3, upsampling.py:142-195) on every level but the last."""

    def __init__(self, cin, cout, n_resnets, add_up, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, 0, eps=eps)
                                      for i in range(n_resnets)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, h):
        for r in self.resnets:
            h = r(h)
        if self.upsamplers:
            h = self.upsamplers[0](h)
        return h


class VaeDecoder(nn.Module):
    """[diffusers autoencoders/vae.Decoder, un-vendored; PARITY UNPINNED] the `vae.decode` of diffusion_feature.py:481-484:
    conv_in -> mid block (resnet, single-head attention, resnet) -> up blocks over reversed(block_out_channels) with
    layers_per_block + 1 resnets each -> GroupNorm + SiLU + conv_out."""

    def __init__(self, block_out=(128, 256, 512, 512), layers=2, latent=4, eps=1e-6):
        super().__init__()
        rev = list(reversed(block_out))
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = VaeMidBlock(rev[0], eps)
        self.up_blocks = nn.ModuleList()
        ch = rev[0]
        for i, co in enumerate(rev):
            self.up_blocks.append(VaeUpBlock(ch, co, layers + 1, i != len(rev) - 1, eps))
            ch = co
        self.conv_norm_out = nn.GroupNorm(32, ch, eps=eps)
        self.conv_out = nn.Conv2d(ch, 3, 3, padding=1)

    def forward(self, z):
        h = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            h = b(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class Vae(nn.Module):
    def __init__(self, scaling_factor, shift_factor=0.0, quant_conv=True, decoder=False, **kw):
        super().__init__()
        self.encoder = VaeEncoder(**kw)
        lat = kw.get("latent", 4)
        if quant_conv:                      # Flux's AutoencoderKL has use_quant_conv False
            self.quant_conv = nn.Conv2d(2 * lat, 2 * lat, 1)
        self.has_quant = quant_conv
        if decoder:                         # only the `vae-out` path needs it (diffusion_feature.py:477-485)
            self.decoder = VaeDecoder(**kw)
            if quant_conv:
                self.post_quant_conv = nn.Conv2d(lat, lat, 1)
        self.scaling_factor = scaling_factor
        self.shift_factor = shift_factor

    def moments(self, x):
        m = self.encoder(x)
        return self.quant_conv(m) if self.has_quant else m

    def decode(self, z):
        """[AutoencoderKL._decode]: post_quant_conv then the decoder."""
        return self.decoder(self.post_quant_conv(z) if self.has_quant else z)


# ------------------------------------------------------------------------------------------ scheduler math
def alphas_cumprod(beta_start=0.00085, beta_end=0.012, n=1000):
    """scaled_linear betas [diffusers schedulers, un-vendored], float32 like diffusers."""
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def resolve_timestep(version, t):
    """set_timesteps(1000) + get_timesteps(1000, t/1000) (diffusion_feature.py:288-295,
    pipeline_pixart_sigma.py:680-714) -> (timestep fed to the denoiser, a, b, input_scale) with
    x_t = a*z + b*eps (scheduler.add_noise) and model input = x_t * input_scale (scale_model_input)."""
    t_start = max(1000 - min(int(1000 * (t / 1000)), 1000), 0)
    ac = alphas_cumprod()
    if version in ("xl", "pgv2"):          # Euler, spacing 'leading', steps_offset 1: timesteps = [1000..1]
        ts = float(1000 - t_start)
        sigma = float(((1 - ac[int(ts)]) / ac[int(ts)]) ** 0.5)
        return ts, 1.0, sigma, 1.0 / math.sqrt(sigma * sigma + 1.0)
    if version == "2-1":                   # Euler built from a PNDM config: spacing 'linspace' -> [999..0]
        ts = float(999 - t_start)
        sigma = float(((1 - ac[int(ts)]) / ac[int(ts)]) ** 0.5)
        return ts, 1.0, sigma, 1.0 / math.sqrt(sigma * sigma + 1.0)
    if version == "1-5":                   # PNDM skip_prk_steps, offset 1: [1000, 999, 999, 998, ..., 1]
        seq = [1000, 999] + list(range(999, 0, -1))
        ts = float(seq[t_start])
        a = float(ac[int(ts)] ** 0.5)
        b = float((1 - ac[int(ts)]) ** 0.5)
        return ts, a, b, 1.0
    if version.startswith("pixart"):
        # [PixArt scheduler_config.json, from memory: DPMSolverMultistepScheduler, beta_schedule 'linear'
        # 1e-4 -> 0.02, timestep_spacing 'linspace', order 1]: timesteps = round(linspace(0, 999, 1001))[::-1][:-1];
        # add_noise = sqrt(abar) z + sqrt(1 - abar) eps (alpha_t / sigma_t of sigma = sqrt((1 - abar) / abar)),
        # scale_model_input = identity. PARITY UNPINNED (un-vendored scheduler).
        import numpy as np
        seq = np.linspace(0, 999, 1001).round()[::-1][:-1]
        ts = float(seq[min(t_start, len(seq) - 1)])
        betas = torch.linspace(0.0001, 0.02, 1000, dtype=torch.float32)
        ac_lin = torch.cumprod(1.0 - betas, dim=0)
        return ts, float(ac_lin[int(ts)] ** 0.5), float((1 - ac_lin[int(ts)]) ** 0.5), 1.0
    raise NotImplementedError(version)


def scheduler_step_coeffs(version, t):
    """`scheduler.step(noise_pred, t, latents)[0]` of diffusion_feature.py:478-480 as prev = c_s * latents + c_m * noise_pred,
    for the first step after set_timesteps(1000) [diffusers 0.32.2 schedulers, un-vendored; PARITY UNPINNED]:
      Euler (xl, pgv2, 2-1; epsilon prediction, s_churn 0): derivative = noise_pred, dt = sigma_next - sigma,
          prev = latents + noise_pred * dt; sigmas = interp(timesteps) ++ [0]
      PNDM (1-5; skip_prk_steps, first step_plms call: counter 0, ets = [noise_pred]): _get_prev_sample with
          prev_timestep = timestep - 1."""
    ts, _, _, _ = resolve_timestep(version, t)
    ts = int(ts)
    ac = alphas_cumprod().double()
    if version in ("xl", "pgv2", "2-1"):
        sig = lambda k: float(((1 - ac[min(k, 999)]) / ac[min(k, 999)]) ** 0.5)
        last = 1 if version != "2-1" else 0          # final timestep of the schedule: the appended sigma 0 follows it
        sigma_next = 0.0 if ts == last else sig(ts - 1)
        return 1.0, sigma_next - sig(ts)
    if version == "1-5":
        a_t = float(ac[min(ts, 999)])
        prev = ts - 1
        a_p = float(ac[prev]) if prev >= 0 else float(ac[0])       # set_alpha_to_one False: final_alpha_cumprod = abar_0
        b_t, b_p = 1.0 - a_t, 1.0 - a_p
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return (a_p / a_t) ** 0.5, -(a_p - a_t) / denom
    raise NotImplementedError("vae-out: scheduler.step of version '%s'" % version)


@torch.no_grad()
def vae_out(version, vae, latents, noise_pred, t):
    """diffusion_feature.py:477-485: latents = scheduler.step(noise_pred, t, latents)[0];
    vae.decode(latents / scaling_factor)."""
    c_s, c_m = scheduler_step_coeffs(version, t)
    prev = c_s * latents + c_m * noise_pred
    return vae.decode(prev / vae.scaling_factor)


def prepare_latents(vae, image, eps_vae, eps_q, a, b):
    """pipeline_pixart_sigma.py:598-677: z = sample(posterior) * scaling_factor; x_t = add_noise(z, eps, t)."""
    m = vae.moments(image.float())
    mean, logvar = m.chunk(2, dim=1)
    logvar = logvar.clamp(-30.0, 20.0)
    z = (mean + torch.exp(0.5 * logvar) * eps_vae) * vae.scaling_factor
    return a * z + b * eps_q


def add_time_ids(img_size):
    """_get_add_time_ids, diffusion_feature.py:534-571 (requires_aesthetics_score False):
    original_size + crops (0,0) + target_size."""
    return torch.tensor([[img_size, img_size, 0, 0, img_size, img_size]], dtype=torch.float32)


@torch.no_grad()
def extract(version, unet, vae, store, image, ctx, pooled, eps_vae, eps_q, t=50, img_size=None):
    """FeatureExtractor.extract (diffusion_feature.py:222-517) for image_type='tensors' already at img_size."""
    store.reset()
    B = image.shape[0]
    ts, a, b, s = resolve_timestep(version, t)
    latents = prepare_latents(vae, image, eps_vae, eps_q, a, b)
    x = latents * s
    ctx_b = ctx.repeat(B, 1, 1) if ctx.shape[0] == 1 else ctx
    kw = {}
    if version in ("xl", "pgv2"):
        kw["text_embeds"] = pooled.repeat(B, 1) if pooled.shape[0] == 1 else pooled
        kw["time_ids"] = add_time_ids(img_size or image.shape[-1]).repeat(B, 1)
    noise_pred = unet(x, ts, ctx_b, **kw)
    return store.stored_feats, latents, noise_pred


@torch.no_grad()
def extract_dit(version, model, vae, store, image, ctx, ctx_mask, eps_vae, eps_q, t=50):
    """FeatureExtractor.extract for the PixArt versions (diffusion_feature.py:277-283,467-474): prompts =
    (embeds, mask, neg_embeds, neg_mask), no repeat over the batch in the reference (it only works at B = 1
    because `timestep` has one entry, attention.py:498-500); here embeds / mask broadcast over B."""
    store.reset()
    B = image.shape[0]
    ts, a, b, s = resolve_timestep(version, t)
    latents = prepare_latents(vae, image, eps_vae, eps_q, a, b)
    ctx_b = ctx.expand(B, -1, -1) if ctx.shape[0] == 1 else ctx
    mask_b = None if ctx_mask is None else (ctx_mask.expand(B, -1) if ctx_mask.shape[0] == 1 else ctx_mask)
    noise_pred = model(latents * s, ts, ctx_b, mask_b)
    return store.stored_feats, latents, noise_pred


# ------------------------------------------------------------------------------------------ stack + correspondence
def resize_concat(feats, out_hw=(128, 128)):
    """aggregation_network.py:62-66 per batch element: interpolate every map bilinearly, cat over channels."""
    return torch.cat([F.interpolate(f, out_hw, mode="bilinear") for f in feats], dim=1)


def points_to_idxs(points, load_size):
    """correspondence_utils.py:140-146 (points in (y, x) order)."""
    import numpy as np
    py = np.clip(points[:, 0], 0, load_size[1] - 1)
    px = np.clip(points[:, 1], 0, load_size[0] - 1)
    return load_size[1] * np.round(py) + np.round(px)


def find_nn_source_correspondences(f1, f2, source_points, load_size):
    """correspondence_utils.py:113-138; returns (points2 (n,2) [y,x], sims (n, hw))."""
    f1 = F.interpolate(f1, load_size, mode="bilinear")
    f2 = F.interpolate(f2, load_size, mode="bilinear")
    idx = torch.from_numpy(points_to_idxs(source_points, load_size)).long()
    b, c = f1.shape[:2]
    f1 = f1.view(b, c, -1).permute(0, 2, 1)[:, idx, :]
    f2 = f2.view(b, c, -1).permute(0, 2, 1)
    f1 = f1 / torch.linalg.norm(f1, dim=-1)[:, :, None]
    f2 = f2 / torch.linalg.norm(f2, dim=-1)[:, :, None]
    sims = torch.matmul(f1, f2.permute(0, 2, 1))
    n = int(math.sqrt(sims.shape[-1]))
    p2 = sims.argmax(dim=-1)
    return torch.stack([p2 // n, p2 % n], dim=-1)[0], sims[0]


# ------------------------------------------------------------------------------------------ segmentation head
def seg_layer_conv_name(layer, model_index=None):
    """segmentation/models/diffusion_segmentor.py:188-192."""
    name = layer.replace("-", "_")
    return name if model_index is None else "%d_%s" % (model_index, name)


def seg_resblock(x, sd, prefix, eps=1e-5):
    """ResBlock.forward, segmentation/models/diffusion_segmentor.py:23-44, inference (eval-mode BatchNorm2d: running
    statistics): x + BN2(conv2(relu(BN1(conv1(x))))). sd holds the module's own parameter names
    (`<prefix>.conv1.0.weight`, `<prefix>.conv1.1.running_mean`, ...)."""
    def conv_bn(h, c):
        h = F.conv2d(h, sd["%s.%s.0.weight" % (prefix, c)], sd["%s.%s.0.bias" % (prefix, c)], stride=1, padding=1)
        return F.batch_norm(h, sd["%s.%s.1.running_mean" % (prefix, c)], sd["%s.%s.1.running_var" % (prefix, c)],
                            sd["%s.%s.1.weight" % (prefix, c)], sd["%s.%s.1.bias" % (prefix, c)], False, 0.0, eps)
    return x + conv_bn(F.relu(conv_bn(x, "conv1")), "conv2")


def seg_extract_feat(features, feature_layers, sd):
    """DiffusionSegmentor.extract_feat, single-extractor branch, segmentation/models/diffusion_segmentor.py:232-246:
    per level, a ResBlock per captured map on its fp32 cast, channel concat, one more ResBlock."""
    outs = []
    for level, res_level in enumerate(feature_layers):
        per = [seg_resblock(features[layer[0]].float(), sd, seg_layer_conv_name(layer[0])) for layer in res_level]
        outs.append(seg_resblock(torch.cat(per, dim=1), sd, seg_layer_conv_name("sum%d" % level)))
    return outs


def seg_extract_feat_multi(features_per_model, feature_layers, c_per_level, sd):
    """DiffusionSegmentor.extract_feat, several-extractors branch, segmentation/models/diffusion_segmentor.py:248-297:
    MultiRes(dim, 4) per map (ONE shared ResBlock applied four times, :46-53), concat, MultiRes(sum, 2) per model and
    level; then the models' results per level are concatenated and go through ResBlock 'amalgemated'."""
    def multires(x, prefix, n):
        for _ in range(n):
            x = seg_resblock(x, sd, prefix + ".res.0")
        return x
    outs = [[] for _ in c_per_level]
    for i, layers in enumerate(feature_layers):
        for level, res_level in enumerate(layers):
            per = [multires(features_per_model[i][layer[0]].float(), seg_layer_conv_name(layer[0], i), 4)
                   for layer in res_level]
            if per:
                outs[level].append(multires(torch.cat(per, dim=1), seg_layer_conv_name("sum%d" % level, i), 2))
    return [seg_resblock(torch.cat(o, dim=1), sd, seg_layer_conv_name("amalgemated", l)) for l, o in enumerate(outs)]
