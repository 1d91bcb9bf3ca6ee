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
"""
import math

import numpy as np


def alphas_cumprod(beta_start=0.00085, beta_end=0.012, n=1000):
    betas = np.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=np.float32) ** 2
    return np.cumprod((1.0 - betas).astype(np.float32), dtype=np.float32)


def resolve(version, t):
    """-> (timestep, a, b, input_scale): x_t = a*z + b*eps, model input = x_t * input_scale."""
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
