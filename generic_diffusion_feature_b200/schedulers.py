"""Timestep / q_sample coefficients the reference obtains from diffusers schedulers.

Reference flow (feature/diffusion_feature.py:288-295): scheduler.set_timesteps(1000); timesteps,_ =
pipe.get_timesteps(1000, t/1000) (pipelines/pixart_alpha/pipeline_pixart_sigma.py:680-714, a copy of the SDXL
img2img one); latent_timestep = timesteps[:1]. Then scheduler.add_noise inside prepare_latents (:673) and
scheduler.scale_model_input (diffusion_feature.py:406). The scheduler classes are un-vendored diffusers 0.32.2
code; their published semantics for the configs models.py selects:

  xl / pgv2 : EulerDiscrete, timestep_spacing 'leading', steps_offset 1 -> timesteps [1000 .. 1]
              add_noise x = z + sigma_t eps, scale_model_input x / sqrt(sigma_t^2 + 1)
  2-1       : EulerDiscrete built from a PNDM config (models.py:38) -> spacing 'linspace' -> [999 .. 0]
  1-5       : PNDM skip_prk_steps, steps_offset 1 -> [1000, 999, 999, 998, ..., 1]
              add_noise sqrt(abar_t) z + sqrt(1 - abar_t) eps, scale_model_input identity
  pixart-*  : DPMSolverMultistep (order 1 for get_timesteps), beta_schedule 'linear' 1e-4 -> 0.02, spacing
              'linspace' -> round(linspace(0, 999, 1001))[::-1][:-1]; add_noise alpha_t z + sigma_t eps with
              alpha_t = sqrt(abar_t), sigma_t = sqrt(1 - abar_t); scale_model_input identity
  flux      : the reference calls the whole img2img pipeline (diffusion_feature.py:246-253) with strength = t/1000,
              guidance_scale 1 and the pipeline's default 28 steps: sigmas = linspace(1, 1/28, 28)
              (pipeline_flux_img2img.py:745), FlowMatchEulerDiscrete with dynamic shifting sigma' = e^mu /
              (e^mu + (1/sigma - 1)), mu = calculate_shift(image_seq_len) (:75-85 with the FLUX.1-dev scheduler
              config base_shift 0.5, max_shift 1.15, seq 256..4096); get_timesteps (:416-425) keeps
              timesteps[int(28 - 28*strength):]; prepare_latents (:565) scale_noise = sigma' eps + (1 - sigma') z;
              the transformer is fed timestep / 1000 = sigma' (:812-815)
"""
import math

import numpy as np


def alphas_cumprod(beta_start=0.00085, beta_end=0.012, n=1000):
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float32) ** 2
    return np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32)


def resolve_flux(t, img_size, num_inference_steps=28, base_shift=0.5, max_shift=1.15, base_seq=256, max_seq=4096):
    """-> (sigma, a, b, 1.0) of the first step the reference's Flux img2img call runs (see module docstring)."""
    strength = t / 1000
    init = min(num_inference_steps * strength, num_inference_steps)
    t_start = int(max(num_inference_steps - init, 0))
    if num_inference_steps - t_start < 1:
        raise ValueError("After adjusting the num_inference_steps by strength parameter: %s, the number of pipeline"
                         "steps is %d which is < 1 and not appropriate for this pipeline."
                         % (strength, num_inference_steps - t_start))     # pipeline_flux_img2img.py:768-772
    sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
    seq_len = (img_size // 8 // 2) ** 2
    m = (max_shift - base_shift) / (max_seq - base_seq)
    mu = seq_len * m + (base_shift - m * base_seq)
    s = float(sigmas[t_start])
    sigma = math.exp(mu) / (math.exp(mu) + (1.0 / s - 1.0))
    sigma = float(np.float32(sigma))
    return sigma, 1.0 - sigma, sigma, 1.0


def resolve(version, t, img_size=None):
    """-> (timestep, a, b, input_scale): x_t = a*z + b*eps, model input = x_t * input_scale."""
    if version == "flux":
        return resolve_flux(t, img_size or 1024)
    init_timestep = min(int(1000 * (t / 1000)), 1000)
    t_start = max(1000 - init_timestep, 0)
    ac = alphas_cumprod()
    if version in ("xl", "pgv2", "2-1"):
        ts = (1000 - t_start) if version != "2-1" else (999 - t_start)
        if ts > 999:
            # index 0 of the 'leading'+offset list is 1000: diffusers interpolates sigma at the table edge
            ts_idx = 999
        else:
            ts_idx = ts
        sigma = math.sqrt((1.0 - float(ac[ts_idx])) / float(ac[ts_idx]))
        return float(ts), 1.0, sigma, 1.0 / math.sqrt(sigma * sigma + 1.0)
    if version == "1-5":
        seq = [1000, 999] + list(range(999, 0, -1))
        ts = seq[min(t_start, len(seq) - 1)]
        a = math.sqrt(float(ac[min(ts, 999)]))
        b = math.sqrt(1.0 - float(ac[min(ts, 999)]))
        return float(ts), a, b, 1.0
    if version.startswith("pixart"):
        seq = np.linspace(0, 999, 1001).round()[::-1][:-1]
        ts = int(seq[min(t_start, len(seq) - 1)])
        betas = np.linspace(0.0001, 0.02, 1000, dtype=np.float32)
        ac_lin = np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32)
        return float(ts), math.sqrt(float(ac_lin[ts])), math.sqrt(1.0 - float(ac_lin[ts])), 1.0
    raise NotImplementedError(version)


def step_coeffs(version, t):
    """`scheduler.step(noise_pred, t, latents)[0]` of diffusion_feature.py:478-480 (the `vae-out` path) as
    prev = c_s * latents + c_m * noise_pred, for the first step after set_timesteps(1000) (module docstring):
      Euler (xl / pgv2 / 2-1, epsilon prediction, s_churn 0): derivative = noise_pred, prev = latents + noise_pred *
          (sigma_next - sigma), sigmas = interp(timesteps) followed by 0
      PNDM (1-5, skip_prk_steps: the first step_plms call, counter 0, ets = [noise_pred]): _get_prev_sample with
          prev_timestep = timestep - 1, final_alpha_cumprod = abar_0 (set_alpha_to_one False)
    PixArt (DPM-Solver on an 8-channel output the reference does not split before step()) and Flux (whole pipeline call)
    have no working `vae-out` in the reference either."""
    ts = int(resolve(version, t)[0])
    ac = alphas_cumprod().astype(np.float64)
    if version in ("xl", "pgv2", "2-1"):
        def sig(k):
            k = min(k, 999)
            return math.sqrt((1.0 - ac[k]) / ac[k])
        last = 1 if version != "2-1" else 0
        sigma_next = 0.0 if ts == last else sig(ts - 1)
        return 1.0, sigma_next - sig(ts)
    if version == "1-5":
        a_t = float(ac[min(ts, 999)])
        prev = ts - 1
        a_p = float(ac[prev]) if prev >= 0 else float(ac[0])
        b_t, b_p = 1.0 - a_t, 1.0 - a_p
        denom = a_t * math.sqrt(b_p) + math.sqrt(a_t * b_t * a_p)
        return math.sqrt(a_p / a_t), -(a_p - a_t) / denom
    raise NotImplementedError("vae-out: scheduler.step for version '%s' is not built (the reference's own vae-out "
                              "path only runs for the UNet families)" % version)
