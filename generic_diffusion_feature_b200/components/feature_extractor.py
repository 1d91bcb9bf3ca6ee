"""Capture plumbing of the B200 path: host-side mirror of feature/components/feature_extractor.py.

In the reference, `prepare_feature_extractor` (feature_extractor.py:92-288) hangs a FeatureGatherer on every
module and `FeatureStore.store` (:31-76) filters / reshapes / copies each activation as the forward runs. Here
the same id grammar is compiled once into a *feature plan* (`gdf_plan`): every requested id becomes a slot of
a preallocated fp16 arena and the producing CUDA kernel writes into it from its epilogue - there are no hooks
and no copies. `FeatureStore` keeps the reference's attribute surface (to_store, accept_all, stored_feats,
reset, store_idx) so that code written against the reference keeps working.
"""
import ctypes
import json

import torch

from .. import _lib
from .._lib import Slot, check


ATTN_MEAN_PREFIX = "#attnmean:"     # internal plan ids: head-mean attention probabilities of one attention module


def _place_of(block):
    """place_in_unet of an attention module: 'down' / 'mid' / 'up' for UNet blocks; every block of a transformer pipe
    is registered with place 'up' (register_attention_store, feature/components/attention.py:567-593)."""
    head = block.split("-")[0]
    return "up" if head == "vit" else head


def attention_mean_ids(cfg, categories, dit_cfg=None, flux_cfg=None):
    """Internal ids of the head-mean maps the reference's AttentionStore would receive for the selected categories
    ('down_cross', 'up_self', ...; feature/components/attention.py:102-118, 531-603), execution order."""
    ids = []
    if flux_cfg is not None:      # one joint attention per block: cross (text keys) first, then self (image keys)
        for k in range(flux_cfg["layers"] + flux_cfg["single_layers"]):
            for kind in ("cross", "self"):
                if "up_%s" % kind in categories:
                    ids.append(ATTN_MEAN_PREFIX + "vit-block%d-%s" % (k, kind))
        return ids
    base = _dit_feature_ids(dit_cfg) if dit_cfg is not None else _unet_feature_ids(cfg)
    for i in base:
        for kind in ("self", "cross"):
            if i.endswith("-%s-q" % kind):
                block = i[:-len("-%s-q" % kind)]
                if "%s_%s" % (_place_of(block), kind) in categories:
                    ids.append(ATTN_MEAN_PREFIX + block + "-" + kind)
    return ids


def aggregate_attention(means, categories, img_size, transformer=False):
    """AttentionStore.aggregate_attention + the assembly of diffusion_feature.py:488-500 on the head-mean maps:
    keep maps with (img/32)^2 <= Nq <= (img/16)^2 (attention.py:111-112, 541; (img/8)^2 for transformer pipes, :568),
    group per category and size, (b (h w) c -> b c h w), average each group, nearest-resize to (img/8, img/8),
    concatenate on the channel axis."""
    import math
    import torch.nn.functional as F
    lo, hi = (img_size // 32) ** 2, (img_size // (8 if transformer else 16)) ** 2
    groups = {c: {} for c in categories}
    for block, kind, t in means:
        cat = "%s_%s" % (_place_of(block), kind)
        if cat not in groups or not (lo <= t.shape[1] <= hi):
            continue
        size = int(math.sqrt(t.shape[1]))
        r = t.reshape(t.shape[0], size, size, t.shape[2]).permute(0, 3, 1, 2)
        groups[cat].setdefault(size, []).append(r)
    parts = []
    for cat in categories:
        for size, maps in groups[cat].items():
            m = torch.stack(maps).float().mean(0).to(torch.float16)
            parts.append(F.interpolate(m, size=(img_size // 8, img_size // 8)))
    return torch.cat(parts, dim=-3)


def _unet_feature_ids(cfg, layers_per_block=2, with_maps=False):
    """Feature ids of a UNet in execution order (same grammar as feature_extractor.py:125-249): the 472 / 165 non-`map`
    ids of feature/configs/config_{xl,15}_full.json, or with_maps=True all 612 / 213 of them, the attention-probability
    maps interleaved where AttnStoreProcessor gathers them (after the module's q / k / v)."""
    ids = ["unet-in", "unet-after-conv-in"]
    bo = cfg["block_out"]
    n = len(bo)
    tags = ("self-q", "self-k", "self-v", "self-map", "cross-q", "cross-map", "ffn-inner", "out") if with_maps else \
        ("self-q", "self-k", "self-v", "cross-q", "ffn-inner", "out")

    def vit(prefix, depth):
        for k in range(depth):
            for tag in tags:
                ids.append("%s-block%d-%s" % (prefix, k, tag))
        ids.append(prefix + "-out")

    for i in range(n):
        for j in range(layers_per_block):
            p = "down-level%d-repeat%d" % (i, j)
            ids += [p + "-res-increment", p + "-res-out"]
            if cfg["down_attn"][i]:
                vit(p + "-vit", cfg["depth"][i])
        if i != n - 1:
            ids.append("down-level%d-downsampler-out" % i)
    ids += ["mid-repeat0-res-increment", "mid-repeat0-res-out"]
    vit("mid-vit", cfg["depth"][-1])
    ids += ["mid-repeat1-res-increment", "mid-repeat1-res-out"]
    for i in range(n):
        li = n - 1 - i
        for j in range(layers_per_block + 1):
            p = "up-level%d-repeat%d" % (i, j)
            ids += [p + "-res-increment", p + "-res-out"]
            if cfg["up_attn"][i]:
                vit(p + "-vit", cfg["depth"][li])
        if i != n - 1:
            ids.append("up-level%d-upsampler-out" % i)
    ids.append("unet-out")
    return ids


def _dit_feature_ids(cfg, with_maps=False):
    """Ids of the PixArt branch of prepare_feature_extractor (feature_extractor.py:259-286) that
    FeatureStore.store keeps, in execution order: per block self-q/k/v, cross-q, ffn-inner, out; with_maps=True adds the
    attention-probability maps where AttnStoreProcessor gathers them (after the module's q / k / v)."""
    ids = []
    tags = ("self-q", "self-k", "self-v", "self-map", "cross-q", "cross-map", "ffn-inner", "out") if with_maps else \
        ("self-q", "self-k", "self-v", "cross-q", "ffn-inner", "out")
    for k in range(cfg["layers"]):
        for tag in tags:
            ids.append("vit-block%d-%s" % (k, tag))
    return ids


def _flux_feature_ids(cfg, with_maps=False):
    """Ids of the Flux branch of prepare_feature_extractor (feature_extractor.py:98-123) in execution order:
    double blocks q, k, v, attn-out (attention_processor.py:2280-2283, 2355-2356), norm-out (transformer_flux.py:
    200-201), ffn-inner (attention.py:1249-1258), out (:210-211); single blocks (numbered after the double
    blocks) q, k, v, attn-out (attention_processor.py:2285-2289, 2358-2360), out (transformer_flux.py:107-108)."""
    ids = []
    # with_maps: FluxAttnStoreProcessor gathers `cross-map` (image queries x text keys) then `self-map` (image x image)
    # right after q / k / v (feature/components/attention.py:494-502)
    maps = ("cross-map", "self-map") if with_maps else ()
    for k in range(cfg["layers"]):
        for tag in ("q", "k", "v") + maps + ("attn-out", "norm-out", "ffn-inner", "out"):
            ids.append("vit-block%d-%s" % (k, tag))
    for k in range(cfg["layers"], cfg["layers"] + cfg["single_layers"]):
        for tag in ("q", "k", "v") + maps + ("attn-out", "out"):
            ids.append("vit-block%d-%s" % (k, tag))
    return ids


class FeatureStore:
    """Mirror of the reference FeatureStore (feature_extractor.py:8-80) backed by the arena plan."""

    def __init__(self, to_store, resize_ratio, train_unet):
        if to_store:
            self.to_store = to_store
            self.accept_all = False
        else:
            self.to_store = {}
            self.accept_all = True
        self.feats = {}
        self.status = "active"
        self.resize_ratio = resize_ratio
        self.train_unet = train_unet
        self.store_idx = None

    def pause(self):
        self.status = "pause"

    def resume(self):
        self.status = "active"

    def reset(self):
        self.feats = {}   # rebinding, like the reference: previously returned dicts stay valid

    @property
    def stored_feats(self):
        return self.feats


def pool_views(lib, feats, ratio):
    """feature_resize (feature_extractor.py:51-53): every captured map -> F.adaptive_avg_pool2d(feat, (h // r, w // r)),
    by gdf_op_avgpool_nhwc on the token-major fp16 storage; the pooled maps live in one new buffer and are returned as
    (B, C, h // r, w // r) views in the same order. (The reference pools in model dtype before the fp16 cast; here the
    fp16 capture is pooled with fp32 sums - one extra rounding.)"""
    if ratio <= 1 or not feats:
        return feats
    total = 0
    plan = []
    for k, v in feats.items():
        B, C, H, W = v.shape
        OH, OW = H // ratio, W // ratio
        if OH < 1 or OW < 1:
            raise ValueError("feature_resize %d is larger than map '%s' (%d x %d)" % (ratio, k, H, W))
        plan.append((k, v, B, C, H, W, OH, OW, total))
        total += (B * OH * OW * C * 2 + 255) // 256 * 256
    dev = next(iter(feats.values())).device
    buf = torch.empty(total, dtype=torch.uint8, device=dev)
    out = {}
    for k, v, B, C, H, W, OH, OW, off in plan:
        assert v.dtype == torch.float16
        if v.is_contiguous() and C > 1:
            # attention-probability map (B, heads, Nq, Nk), stored like the reference's 4-D tensor: pooled over its
            # last two axes like any other 4-D feature (FeatureStore.store does not special-case it) - one channel
            dst = buf[off:off + B * OH * OW * C * 2].view(torch.float16).view(B, C, OH, OW)
            check(lib.gdf_op_avgpool_nhwc(_lib.ptr(v), _lib.ptr(dst), B * C, H, W, 1, OH, OW, _lib.stream_ptr()))
            out[k] = dst
            continue
        src = v.permute(0, 2, 3, 1)                       # the arena storage: (B, H, W, C) contiguous
        assert src.is_contiguous()
        dst = buf[off:off + B * OH * OW * C * 2].view(torch.float16).view(B, OH, OW, C)
        check(lib.gdf_op_avgpool_nhwc(_lib.ptr(src), _lib.ptr(dst), B, H, W, C, OH, OW, _lib.stream_ptr()))
        out[k] = dst.permute(0, 3, 1, 2)
    return out


class FeaturePlan:
    """Compiled selection: ids -> arena slots for one (batch, img_size)."""

    def __init__(self, pipe, ids, batch, img_size):
        self.ids = list(ids)
        self.batch = batch
        self.img_size = img_size
        lib = pipe.lib
        n = len(self.ids)
        c_ids = (ctypes.c_char_p * max(n, 1))(*[i.encode() for i in self.ids])
        slots = (Slot * max(n, 1))()
        arena_bytes = ctypes.c_int64(0)
        with torch.cuda.device(pipe.dev_index):
            check(lib.gdf_plan(pipe.handle, c_ids, n, batch, img_size, slots, ctypes.byref(arena_bytes)))
        self.arena_bytes = int(arena_bytes.value)
        self.slots = [(self.ids[i], int(slots[i].offset_bytes), slots[i].channels, slots[i].height, slots[i].width,
                       slots[i].order) for i in range(n)]
        self.launches = int(lib.gdf_num_launches(pipe.handle))
        self.workspace_bytes = int(lib.gdf_workspace_bytes(pipe.handle))
        # the handle holds ONE compiled plan: whoever plans next on the same pipe (a second extractor built with
        # external_model=pipe) replaces it; `current()` tells the owner to plan again before replaying
        self.generation = int(lib.gdf_plan_generation(pipe.handle))
        self._pipe = pipe

    def current(self):
        return int(self._pipe.lib.gdf_plan_generation(self._pipe.handle)) == self.generation

    def views(self, arena):
        """dict id -> fp16 (B, C, h, w) view of the arena, insertion order = execution order (the reference's
        dict order). Token-major storage: like the reference's ViT features (einops view of (B, N, C),
        feature_extractor.py:46-48) the (B,C,h,w) tensor has channel stride 1."""
        out = []
        seen = set()
        for fid, off, C, H, W, order in self.slots:
            if off < 0 or fid in seen or fid.startswith(ATTN_MEAN_PREFIX):
                continue
            seen.add(fid)
            nbytes = self.batch * H * W * C * 2
            if fid.endswith("-map"):
                # attention probabilities: (B, heads, Nq, Nk) contiguous, exactly the 4-D tensor the reference stores
                t = arena[off:off + nbytes].view(torch.float16).view(self.batch, C, H, W)
            else:
                t = arena[off:off + nbytes].view(torch.float16).view(self.batch, H, W, C).permute(0, 3, 1, 2)
            out.append((order, fid, t))
        out.sort(key=lambda x: x[0])
        return {fid: t for _, fid, t in out}

    def attention_means(self, arena):
        """Internal head-mean attention maps, execution order: [(block id, 'self'|'cross', (B, Nq, Nk) fp16)]."""
        out = []
        for fid, off, C, H, W, order in self.slots:
            if off < 0 or not fid.startswith(ATTN_MEAN_PREFIX):
                continue
            block, kind = fid[len(ATTN_MEAN_PREFIX):].rsplit("-", 1)
            t = arena[off:off + self.batch * H * W * 2].view(torch.float16).view(self.batch, H, W)
            out.append((order, block, kind, t))
        out.sort(key=lambda x: x[0])
        return [(b, k, t) for _, b, k, t in out]


def prepare_feature_extractor(version, pipe, config, resize_ratio, train_unet):
    """Same signature / return type as the reference (feature_extractor.py:92): config is a JSON path, a dict
    {feature_id: bool} or None/{} for accept-all."""
    if isinstance(config, str):
        with open(config, "r") as f:
            config = json.load(f)
    if train_unet:
        raise NotImplementedError("train_unet needs autograd through the kernels (SURVEY.md 8f, not built)")
    if not isinstance(resize_ratio, int) or resize_ratio < 1:
        raise ValueError("feature_resize must be a positive integer")
    return FeatureStore(config, resize_ratio, train_unet)


def selected_ids(feature_store, pipe):
    """Ids to plan: enabled JSON keys in file order, or every id of the architecture when accept_all
    (feature_extractor.py:10-15,36). `map` ids need the attention-probability path and raise."""
    if feature_store.accept_all:
        if getattr(pipe, "flux_cfg", None):
            return _flux_feature_ids(pipe.flux_cfg, with_maps=True)
        if getattr(pipe, "dit_cfg", None):
            return _dit_feature_ids(pipe.dit_cfg, with_maps=True)   # accept-all installs the storing processors too
        # an empty config makes the reference install its storing attention processors (diffusion_feature.py:74-77), so
        # accept-all includes every `...-map` (this is how config_{xl,15}_full.json were produced, :502-514)
        return _unet_feature_ids(pipe.unet_cfg, with_maps=True)
    ids = [k for k, v in feature_store.to_store.items() if v]
    for k in ids:
        if "map" in k:
            continue          # per-layer attention probabilities (slow materialising path, like the reference's)
        if k == "attn":
            raise NotImplementedError("feature id 'attn' comes from FeatureExtractor(attention=[...])")
    # `vae-out` is not a capture site: FeatureExtractor.extract decodes it after the forward (diffusion_feature.py:477-485)
    return [k for k in ids if k != "vae-out"]
