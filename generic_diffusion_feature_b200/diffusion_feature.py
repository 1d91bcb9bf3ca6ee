"""Drop-in `FeatureExtractor` for the B200 path (mirror of feature/diffusion_feature.py in the reference).

Same constructor and method signatures (diffusion_feature.py:26-40,118-144,149-235,520-527). Underneath, the
diffusers pipeline is replaced by libgdf_b200.so: `extract` = gdf_encode_noise (VAE encode + posterior sample
+ q_sample + scale_model_input) followed by gdf_denoise_capture (one UNet forward whose kernels write the
selected activations straight into a preallocated fp16 arena). There is no PyTorch compute on the path and no
CPU fallback.

Differences a user can observe (all documented in INTEGRATION.md):
  * returned maps are fp16 CUDA tensors shaped (B, C, h, w) like the reference's, but ALL of them are
    token-major-strided views of one arena (the reference returns conv maps NCHW-contiguous and ViT maps
    token-major, feature_extractor.py:46-48); values and indexing are identical, `.contiguous()` restores NCHW;
  * `extract(..., noise=(eps_vae, eps_q))` optionally injects the two Gaussian draws the reference makes
    un-seeded (pipeline_pixart_sigma.py:644,671) so that results are reproducible;
  * the PixArt versions take prompts = (embeds, mask, neg_embeds, neg_mask) like the reference
    (diffusion_feature.py:277-283) and, unlike it (B = 1 only, attention.py:498-500), any batch size;
  * version 'flux' (diffusion_feature.py:246-253 calls the whole FluxImg2ImgPipeline with strength = t/1000 and
    guidance_scale 1): prompts = (t5_embeds (1|B, 512, 4096), pooled_clip (1|B, 768)) as returned by encode_prompt
    here (the reference passes raw strings to the pipeline); noise = (eps_vae, eps_q) with 16 latent channels;
  * weights come from `external_model`, else from the diffusers-layout directory in GDF_MODEL_DIR[/<version>], else - only
    with GDF_SYNTHETIC=1 - from the deterministic random initialisation; there is no silent default (models.py here);
  * control / train_unet / denoising_from / DDIM inversion, Hunyuan and IF raise NotImplementedError.
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, schedulers
from ._lib import check
from .components.feature_extractor import (FeaturePlan, aggregate_attention, attention_mean_ids, pool_views,
                                           prepare_feature_extractor, selected_ids)
from .components.models import get_diffusion_model


class FeatureExtractor(nn.Module):
    def __init__(self,
                 layer,  # the filename of layer json or a pre-loaded dict
                 version,  # xl, pgv2, 1-5, 2-1
                 device,
                 dtype='float16',
                 img_size=1024,  # 512 for 1-5, 1024 otherwise
                 offline_lora=None,
                 offline_lora_filename=None,
                 feature_resize=1,
                 control=None,
                 attention=None,
                 train_unet=False,
                 external_model=None,
                 ):
        super().__init__()
        if control:
            # components/controlnet.py runs diffusers ControlNetModel checkpoints (+ cv2 / Midas preprocessors) to produce
            # the residuals; that encoder network is not built here. Its OUTPUTS are taken: extract(control_residuals=...)
            raise NotImplementedError("the ControlNet encoder (feature/components/controlnet.py) is not built on the B200 "
                                      "path; pass its outputs to extract(..., control_residuals=(down, mid)) instead")
        if external_model:
            pipe = external_model            # diffusion_feature.py:46-47: the seam for pre-built pipes
        else:
            # the VAE decoder is only loaded when the layer selection asks for `vae-out` (diffusion_feature.py:477-485)
            try:
                import json
                sel = layer if isinstance(layer, dict) else (json.load(open(layer)) if layer else {})
                wants_decoder = bool(sel.get('vae-out', False))
            except (OSError, ValueError, AttributeError):
                wants_decoder = False
            pipe = get_diffusion_model(version, dtype, offline_lora, offline_lora_filename, device=device,
                                       with_decoder=wants_decoder)
        self.feature_store = prepare_feature_extractor(version, pipe, layer, feature_resize, train_unet)
        self.store_vae_output = bool(self.feature_store.to_store.get('vae-out', False))
        if self.store_vae_output:
            # diffusion_feature.py:477-485: scheduler.step + vae.decode. UNet families only (the reference's own vae-out
            # does not run for the transformer pipes either: an 8-channel PixArt output goes unsplit into step(), the
            # Flux branch returns before it), and the pipe must have been given the VAE decoder weights.
            if getattr(pipe, "dit_cfg", None) is not None or getattr(pipe, "flux_cfg", None) is not None:
                raise NotImplementedError("vae-out: only the UNet families (xl, pgv2, 2-1, 1-5) have a scheduler.step + "
                                          "vae.decode path")
            schedulers.step_coeffs(version, 50)      # raises for a version without a restated scheduler.step
            if not getattr(pipe, "has_decoder", False):
                raise ValueError("vae-out needs the VAE decoder: load a state dict / model_dir that holds "
                                 "'vae.decoder.*' and 'vae.post_quant_conv.*' (load_diffusers_dir(..., "
                                 "with_decoder=True) / synthetic_state_dict(..., with_decoder=True))")
        self.pipe = pipe
        self.control_pipe = None
        self.attention_store = None
        self.version = version
        self.img_size = img_size
        self.device = device
        self.control = control
        self.attention = attention
        self._plan = None
        self._plan_ctx_len = None
        self._copy_stream = None
        self._ids = selected_ids(self.feature_store, pipe)
        if attention:
            # register_attention_store (diffusion_feature.py:67-68): the head-mean probabilities of every attention
            # module of the selected categories become internal plan slots; `extract` aggregates them into `attn`
            self._ids = list(self._ids) + attention_mean_ids(getattr(pipe, "unet_cfg", None), list(attention),
                                                             dit_cfg=getattr(pipe, "dit_cfg", None),
                                                             flux_cfg=getattr(pipe, "flux_cfg", None))

    # ------------------------------------------------------------------------------------------ images
    def _preprocess_basic(self, x):
        return x.resize((self.img_size, self.img_size)).convert("RGB")

    def preprocess_image(self, x, is_tensor=False):
        """PIL -> (1,3,S,S) in [-1,1] (VaeImageProcessor.preprocess semantics: /255, NCHW, 2x-1);
        tensors pass through (diffusion_feature.py:122-140)."""
        if is_tensor:
            return x
        arr = np.asarray(self._preprocess_basic(x), dtype=np.float32) / 255.0
        return torch.from_numpy(arr).permute(2, 0, 1)[None] * 2.0 - 1.0

    def restore_from_tensor_to_image(self, x):
        from PIL import Image
        x = ((x.detach().float().cpu() / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)
        return [Image.fromarray(i.permute(1, 2, 0).numpy()) for i in x]

    def stage_images(self, images_host):
        """Start the host -> device copy of a pinned (B, 3, S, S) float32 batch on a side stream and return the device
        tensor; `extract(image=<that tensor>, image_type='tensors')` makes its stream wait for the copy. Lets a caller
        that loops over batches overlap the copy of batch i + 1 with the forward of batch i (what the CLI's prefetch
        thread does, extract_feature.py). Extension: the reference copies synchronously inside `extract`."""
        dev = torch.device(self.pipe.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self._copy_stream):
            t = images_host.to(dev, torch.float32, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        t._gdf_ready = ev
        return t

    # ------------------------------------------------------------------------------------------ prompts
    def encode_prompt(self, prompt_str=None, prompt_file=None):
        """The reference encodes text with the pipeline's CLIP encoders (diffusion_feature.py:149-206). No
        text-encoder weights exist offline, and text encoding is outside the extraction path (empty-prompt
        conditioning, encoded once): this returns deterministic stand-in embeddings of the right shapes
        (N(0,1), seeded by the prompt text). Real embeddings can be passed to `extract` directly."""
        assert prompt_str != None and prompt_file == None or prompt_str == None and prompt_file != None
        if prompt_file:
            with open(prompt_file, 'r') as f:
                prompt_str = f.read()
        import warnings
        import zlib
        warnings.warn("FeatureExtractor.encode_prompt on the B200 path returns STAND-IN embeddings (N(0,1), seeded by "
                      "the prompt text): no text encoder is part of this path. Encode the prompt once with the "
                      "reference and pass the embeddings to extract() for real conditioning.", stacklevel=2)
        g = torch.Generator().manual_seed(zlib.crc32(prompt_str.encode()))
        if getattr(self.pipe, "flux_cfg", None):
            # Flux: T5 sequence (max_sequence_length 512) + CLIP pooled vector (pipeline_flux_img2img.py encode_prompt)
            fc = self.pipe.flux_cfg
            return torch.randn(1, fc["ctx_len"], fc["joint_dim"], generator=g), torch.randn(1, fc["pooled_dim"], generator=g)
        if getattr(self.pipe, "dit_cfg", None):
            # PixArt: (prompt_embeds, prompt_attention_mask, negative_embeds, negative_mask), T5 length 300 for
            # Sigma (diffusion_feature.py:193-205); the stand-in mask keeps every token
            cd = self.pipe.dit_cfg["caption_dim"]
            n_tok = self.pipe.dit_cfg.get("prompt_len", 300)     # 120 T5 tokens for PixArt-alpha, 300 for Sigma
            emb = torch.randn(1, n_tok, cd, generator=g)
            neg = torch.randn(1, n_tok, cd, generator=g)
            return emb, torch.ones(1, n_tok), neg, torch.ones(1, n_tok)
        ctx_dim = self.pipe.unet_cfg["ctx_dim"]
        emb = torch.randn(1, 77, ctx_dim, generator=g)
        neg = torch.randn(1, 77, ctx_dim, generator=g)
        if self.version in ('xl', 'pgv2'):
            uc = self.pipe.unet_cfg
            pd = uc["add_in"] - 6 * uc["add_time_dim"]      # 1280 for SDXL (projection_dim of text_encoder_2)
            pooled = torch.randn(1, pd, generator=g)
            npooled = torch.randn(1, pd, generator=g)
            return emb, neg, pooled, npooled
        return emb, neg, None, None

    def offload_prompt_encoder(self, persistent=False):
        return None   # no text encoder is resident on the B200 path

    # ------------------------------------------------------------------------------------------ extract
    def _ensure_plan(self, batch_size, ctx_len):
        if (self._plan is None or self._plan.batch != batch_size or self._plan.img_size != self.img_size
                or self._plan_ctx_len != ctx_len or not self._plan.current()):
            check(self.pipe.lib.gdf_set_ctx_len(self.pipe.handle, ctx_len))
            self._plan = FeaturePlan(self.pipe, self._ids, batch_size, self.img_size)
            self._plan_ctx_len = ctx_len
        return self._plan

    @torch.no_grad()
    def extract(self,
                prompts,  # the same as the outputs of the last function
                batch_size,
                image,
                image_type='image',  # otherwise: tensors
                t=50,
                denoising_from=None,
                use_control=False,
                use_ddim_inversion=False,
                noise=None,   # extension: (eps_vae, eps_q), each (B,4,S/8,S/8)
                control_residuals=None,   # extension: (down_block_res_samples, mid_block_res_sample) of a ControlNet
                ):
        if denoising_from:
            raise NotImplementedError("denoising_from (multi-step denoise loop) is deprecated in the reference and "
                                      "not built here")
        if use_ddim_inversion:
            raise NotImplementedError("DDIM inversion is outside the B200 hot path")
        if use_control and control_residuals is None:
            raise NotImplementedError("use_control needs the ControlNet outputs: pass control_residuals=(down, mid) "
                                      "(unet_2d_condition.py:1236-1275); the ControlNet encoder itself is not built here")
        if self.feature_store.store_idx is not None:
            raise NotImplementedError("background extraction hooks a generation loop, which is not built here")
        pipe = self.pipe
        dev = pipe.device
        self.feature_store.reset()
        is_dit = getattr(pipe, "dit_cfg", None) is not None
        is_flux = getattr(pipe, "flux_cfg", None) is not None
        ctx_mask = None
        if is_flux:
            prompt_embeds, pooled = prompts[0], prompts[1]
        elif is_dit:
            prompt_embeds, ctx_mask, _neg, _neg_mask = prompts      # diffusion_feature.py:277-283
            pooled = None
            if ctx_mask is not None and ctx_mask.shape[0] == 1:
                ctx_mask = ctx_mask.repeat(batch_size, 1)
        else:
            prompt_embeds, _neg, pooled, _npooled = prompts
        prompt_embeds = prompt_embeds.repeat(batch_size, 1, 1) if prompt_embeds.shape[0] == 1 else prompt_embeds
        if pooled is not None and pooled.shape[0] == 1:
            pooled = pooled.repeat(batch_size, 1)
        timestep, qa, qb, qs = schedulers.resolve("flux" if is_flux else self.version, t, self.img_size)
        # 6. prepare image (diffusion_feature.py:357-364)
        if image_type == 'image':
            image = torch.concat([self.preprocess_image(r) for r in image], dim=0)
        ready = getattr(image, "_gdf_ready", None)      # staged by stage_images(): the copy runs on a side stream
        image = image.to(dev, torch.float32, non_blocking=True)
        if ready is not None:
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ready)
            image.record_stream(cur)
        lat_ch = pipe.vae_cfg["latent"]
        is_latents = image.shape[1] == lat_ch  # prepare_latents: latent-channel input is taken as latents (:623-624)
        if not is_latents and image.shape[-2:] != (self.img_size, self.img_size):
            image = F.interpolate(image, (self.img_size, self.img_size), mode='bilinear')
        image = image.contiguous()
        if image.shape[0] != batch_size:
            raise ValueError("got %d images for batch_size %d" % (image.shape[0], batch_size))
        plan = self._ensure_plan(batch_size, int(prompt_embeds.shape[1]))
        L = self.img_size // 8
        if noise is None:
            eps_vae = torch.randn(batch_size, lat_ch, L, L, device=dev)
            eps_q = torch.randn(batch_size, lat_ch, L, L, device=dev)
        else:
            eps_vae = noise[0].to(dev, torch.float32).contiguous()
            eps_q = noise[1].to(dev, torch.float32).contiguous()
        ctx = prompt_embeds.to(dev, torch.float32).contiguous()
        pooled_d = pooled.to(dev, torch.float32).contiguous() if pooled is not None else None
        # the C ABI takes raw pointers: every extent is checked here against the architecture first
        ctx_dim = (pipe.flux_cfg["joint_dim"] if is_flux else pipe.dit_cfg["caption_dim"] if is_dit
                   else pipe.unet_cfg["ctx_dim"])
        if ctx.dim() != 3 or ctx.shape[0] != batch_size or ctx.shape[2] != ctx_dim:
            raise ValueError("prompt embeddings must be (%d | 1, L, %d) for version '%s', got %s"
                             % (batch_size, ctx_dim, self.version, tuple(prompt_embeds.shape)))
        pooled_dim = (pipe.flux_cfg["pooled_dim"] if is_flux else
                      (pipe.unet_cfg["add_in"] - 6 * pipe.unet_cfg["add_time_dim"]) if (not is_dit and
                                                                                        pipe.unet_cfg["add_time_dim"])
                      else None)
        if pooled_dim is not None:
            if pooled_d is None or tuple(pooled_d.shape) != (batch_size, pooled_dim):
                raise ValueError("pooled text embeddings must be (%d | 1, %d) for version '%s', got %s"
                                 % (batch_size, pooled_dim, self.version,
                                    None if pooled is None else tuple(pooled.shape)))
        for name, e_ in (("eps_vae", eps_vae), ("eps_q", eps_q)):
            if tuple(e_.shape) != (batch_size, lat_ch, L, L):
                raise ValueError("noise[%s] must be (%d, %d, %d, %d), got %s"
                                 % (name, batch_size, lat_ch, L, L, tuple(e_.shape)))
        if is_latents and tuple(image.shape) != (batch_size, lat_ch, L, L):
            raise ValueError("latent input must be (%d, %d, %d, %d), got %s"
                             % (batch_size, lat_ch, L, L, tuple(image.shape)))
        if ctx_mask is not None and tuple(ctx_mask.shape) != (batch_size, ctx.shape[1]):
            raise ValueError("prompt attention mask must be (%d | 1, %d), got %s"
                             % (batch_size, ctx.shape[1], tuple(ctx_mask.shape)))
        time_ids = None
        if self.version in ('xl', 'pgv2'):
            # _get_add_time_ids (diffusion_feature.py:534-571): original_size + crop (0,0) + target_size
            s = float(self.img_size)
            time_ids = torch.tensor([[s, s, 0.0, 0.0, s, s]], device=dev).repeat(batch_size, 1).contiguous()
        arena = torch.empty(plan.arena_bytes, dtype=torch.uint8, device=dev)
        lib = pipe.lib
        # vae-out: the noised latents (before input scaling) and the noise prediction leave the two passes as fp32 NCHW
        lat_out = torch.empty(batch_size, lat_ch, L, L, device=dev) if self.store_vae_output else None
        npred_out = torch.empty(batch_size, lat_ch, L, L, device=dev) if self.store_vae_output else None
        with torch.cuda.device(pipe.dev_index):
            st = _lib.stream_ptr()
            if is_latents:
                check(lib.gdf_encode_latents(pipe.handle, _lib.ptr(image), _lib.ptr(eps_q), qa, qb, qs,
                                             _lib.ptr(lat_out), st))
            else:
                check(lib.gdf_encode_noise(pipe.handle, _lib.ptr(image), _lib.ptr(eps_vae), _lib.ptr(eps_q), qa, qb,
                                           qs, _lib.ptr(lat_out), st))
            if is_flux:
                mask_d = None
                key = (ctx.shape[1], L // 2)
                if getattr(self, "_rope_key", None) != key:
                    from .components.models import flux_rope_tables
                    cos, sin = flux_rope_tables(pipe.flux_cfg, ctx.shape[1], L // 2)
                    self._rope = (cos.to(dev).contiguous(), sin.to(dev).contiguous())
                    self._rope_key = key
                check(lib.gdf_denoise_capture_flux(pipe.handle, timestep, 1.0, _lib.ptr(ctx), ctx.shape[1],
                                                   _lib.ptr(pooled_d), _lib.ptr(self._rope[0]), _lib.ptr(self._rope[1]),
                                                   _lib.ptr(arena), arena.numel(), None, st))
            elif is_dit:
                mask_d = ctx_mask.to(dev, torch.float32).contiguous() if ctx_mask is not None else None
                check(lib.gdf_denoise_capture_dit(pipe.handle, timestep, _lib.ptr(ctx), ctx.shape[1], _lib.ptr(mask_d),
                                                  _lib.ptr(arena), arena.numel(), None, st))
            else:
                mask_d = None
                ctrl_keep = self._set_control_residuals(control_residuals, batch_size)
                check(lib.gdf_denoise_capture(pipe.handle, timestep, _lib.ptr(ctx), ctx.shape[1], _lib.ptr(pooled_d),
                                              _lib.ptr(time_ids), _lib.ptr(arena), arena.numel(), _lib.ptr(npred_out),
                                              st))
                if self.store_vae_output:
                    # latents = scheduler.step(noise_pred, t, latents)[0]; vae.decode(latents / scaling_factor)
                    # (diffusion_feature.py:478-484): both steps are linear in (latents, noise_pred)
                    c_s, c_m = schedulers.step_coeffs(self.version, t)
                    sf = pipe.vae_cfg["scaling_factor"]
                    vae_img = torch.empty(batch_size, self.img_size, self.img_size, 3, device=dev)
                    check(lib.gdf_plan_decoder(pipe.handle))
                    check(lib.gdf_decode_latents(pipe.handle, _lib.ptr(lat_out), c_s / sf, _lib.ptr(npred_out), c_m / sf,
                                                 _lib.ptr(vae_img), st))
        feats = plan.views(arena)
        if self.feature_store.resize_ratio > 1:                  # feature_extractor.py:51-53
            with torch.cuda.device(pipe.dev_index):
                feats = pool_views(lib, feats, self.feature_store.resize_ratio)
        if self.store_vae_output:      # stored as the pipe's dtype, outside FeatureStore.store (diffusion_feature.py:485)
            feats['vae-out'] = vae_img.permute(0, 3, 1, 2).to(torch.float16)
        if self.attention:                                       # diffusion_feature.py:488-500
            feats['attn'] = aggregate_attention(plan.attention_means(arena), list(self.attention), self.img_size,
                                                transformer=(getattr(self.pipe, 'dit_cfg', None) is not None or
                                                             getattr(self.pipe, 'flux_cfg', None) is not None))
        if self.feature_store.accept_all:
            feats = {k: v.cpu() for k, v in feats.items()}   # feature_extractor.py:65-66
        self.feature_store.feats = feats
        # keep inputs alive until the stream has consumed them
        self._keepalive = (image, eps_vae, eps_q, ctx, pooled_d, time_ids, arena, mask_d,
                           locals().get("ctrl_keep"), lat_out, npred_out, locals().get("vae_img"))
        return self.feature_store.stored_feats

    def _set_control_residuals(self, control_residuals, batch_size):
        """ControlNet outputs -> gdf_set_control_residuals (UNet families): `down` = one (B, C_i, s_i, s_i) tensor per skip
        in push order, `mid` = (B, C, s, s); None clears them. Shapes are checked against the planned UNet."""
        pipe, lib = self.pipe, self.pipe.lib
        if control_residuals is None:
            check(lib.gdf_set_control_residuals(pipe.handle, None, 0, None))
            return None
        if getattr(pipe, "unet_cfg", None) is None:
            raise NotImplementedError("control residuals are inputs of the UNet families")
        down, mid = control_residuals
        ch = (ctypes.c_int * 32)()
        sd = (ctypes.c_int * 32)()
        n = lib.gdf_control_residual_shapes(pipe.handle, ch, sd, 32)
        if n < 0:
            check(n)
        if len(down) != n:
            raise ValueError("got %d down residuals, this UNet has %d skip tensors" % (len(down), n))
        dev = pipe.device
        keep = []
        for i, d_ in enumerate(down):
            if tuple(d_.shape) != (batch_size, ch[i], sd[i], sd[i]):
                raise ValueError("down residual %d must be %s, got %s" % (i, (batch_size, ch[i], sd[i], sd[i]),
                                                                       tuple(d_.shape)))
            keep.append(d_.to(dev, torch.float32).contiguous())
        mid_d = None
        if mid is not None:
            if tuple(mid.shape) != (batch_size, ch[n - 1], sd[n - 1], sd[n - 1]):
                raise ValueError("mid residual must be %s, got %s" % ((batch_size, ch[n - 1], sd[n - 1], sd[n - 1]),
                                                                    tuple(mid.shape)))
            mid_d = mid.to(dev, torch.float32).contiguous()
            keep.append(mid_d)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in keep[:n]])
        check(lib.gdf_set_control_residuals(pipe.handle, ptrs, n, _lib.ptr(mid_d)))
        return keep

    def set_background_extraction(self, idxs):
        self.feature_store.store_idx = idxs

    def get_background_extraction(self):
        return {k: v['feat'] for k, v in self.feature_store.feats.items()}
