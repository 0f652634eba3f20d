"""Mirror of image_model/diffusion/__init__.py: the `create_diffusion` factory (same signature and defaults)."""
from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps

__all__ = ["create_diffusion", "SpacedDiffusion", "space_timesteps", "gd"]


def create_diffusion(timestep_respacing, noise_schedule="linear", use_kl=False, sigma_small=True, predict_xstart=True,
                     learn_sigma=False, rescale_learned_sigmas=False, diffusion_steps=1000):
    """diffusion/__init__.py:10-46: linear 1000-step schedule, START_X prediction, fixed-small variance, MSE loss by
    default; `timestep_respacing` "" / None keeps every step, "250" keeps 250 evenly spaced ones."""
    betas = gd.get_named_beta_schedule(noise_schedule, diffusion_steps)
    if use_kl:
        loss_type = gd.LossType.RESCALED_KL
    elif rescale_learned_sigmas:
        loss_type = gd.LossType.RESCALED_MSE
    else:
        loss_type = gd.LossType.MSE
    if timestep_respacing is None or timestep_respacing == "":
        timestep_respacing = [diffusion_steps]
    if learn_sigma:
        var_type = gd.ModelVarType.LEARNED_RANGE
    else:
        var_type = gd.ModelVarType.FIXED_SMALL if sigma_small else gd.ModelVarType.FIXED_LARGE
    return SpacedDiffusion(
        use_timesteps=space_timesteps(diffusion_steps, timestep_respacing),
        betas=betas,
        model_mean_type=gd.ModelMeanType.START_X if predict_xstart else gd.ModelMeanType.EPSILON,
        model_var_type=var_type,
        loss_type=loss_type,
    )
