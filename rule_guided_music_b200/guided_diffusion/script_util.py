"""create_diffusion and the argparse helpers -- mirror of guided_diffusion/script_util.py of the reference
(diffusion_defaults :13-26, create_diffusion :106-126, create_gaussian_diffusion :462-500, add_dict_to_argparser
:503-513, args_to_dict :516-517, str2bool :520-531).  The UNet / classifier / super-resolution factories are not on
the sampling path: `create_model_and_diffusion` (the UNet factory, script_util.py:129-186) raises; `NUM_CLASSES` and
`model_and_diffusion_defaults` (:10, :74-97) are kept because scripts/sample_rule.py imports them and feeds the
defaults to its argument parser (:311)."""
import argparse

NUM_CLASSES = 3  # number of datasets (script_util.py:10)

from . import gaussian_diffusion as gd
from .respace import SpacedDiffusion, space_timesteps


def diffusion_defaults():
    return dict(learn_sigma=False, diffusion_steps=1000, noise_schedule="linear", timestep_respacing="", use_kl=False,
                predict_xstart=False, rescale_timesteps=False, rescale_learned_sigmas=False)


def model_and_diffusion_defaults():
    """Argument defaults of the sampling scripts (script_util.py:74-97): the UNet fields are parsed and ignored by the
    DiT path, the diffusion fields are what create_diffusion takes."""
    res = dict(image_size=128, in_channels=1, num_channels=128, num_res_blocks=2, num_heads=4, num_heads_upsample=-1,
               num_head_channels=-1, attention_resolutions="32,16,8", channel_mult="", dropout=0.0, class_cond=False,
               use_checkpoint=False, use_scale_shift_norm=True, resblock_updown=False, use_fp16=False,
               use_new_attention_order=False)
    res.update(diffusion_defaults())
    return res


def create_model_and_diffusion(*args, **kwargs):
    raise NotImplementedError("create_model_and_diffusion builds the UNet denoiser, which is not on the B200 sampling "
                              "path: use DiT_models[name](...) and create_diffusion(...) as scripts/sample_rule.py does")


def create_diffusion(learn_sigma=False, diffusion_steps=1000, noise_schedule="linear", timestep_respacing="",
                     use_kl=False, predict_xstart=False, rescale_timesteps=False, rescale_learned_sigmas=False):
    return create_gaussian_diffusion(steps=diffusion_steps, learn_sigma=learn_sigma, noise_schedule=noise_schedule,
                                     use_kl=use_kl, predict_xstart=predict_xstart,
                                     rescale_timesteps=rescale_timesteps,
                                     rescale_learned_sigmas=rescale_learned_sigmas,
                                     timestep_respacing=timestep_respacing)


def create_gaussian_diffusion(*, steps=1000, learn_sigma=False, sigma_small=False, noise_schedule="linear",
                              use_kl=False, predict_xstart=False, rescale_timesteps=False,
                              rescale_learned_sigmas=False, timestep_respacing=""):
    betas = gd.get_named_beta_schedule(noise_schedule, steps)
    if use_kl:
        loss_type = gd.LossType.RESCALED_KL
    elif rescale_learned_sigmas:
        loss_type = gd.LossType.RESCALED_MSE
    else:
        loss_type = gd.LossType.MSE
    if learn_sigma:
        var_type = gd.ModelVarType.LEARNED_RANGE
    else:
        var_type = gd.ModelVarType.FIXED_SMALL if sigma_small else gd.ModelVarType.FIXED_LARGE
    return SpacedDiffusion(
        use_timesteps=space_timesteps(steps, timestep_respacing or [steps]),
        betas=betas,
        model_mean_type=gd.ModelMeanType.START_X if predict_xstart else gd.ModelMeanType.EPSILON,
        model_var_type=var_type,
        loss_type=loss_type,
        rescale_timesteps=rescale_timesteps,
    )


def add_dict_to_argparser(parser, default_dict):
    for k, v in default_dict.items():
        v_type = type(v)
        if v is None:
            v_type = str
        elif isinstance(v, bool):
            v_type = str2bool
        if k == "image_size":  # `--image_size 128 16`: scripts index args.image_size[0] / [1] (reference :510-511)
            parser.add_argument(f"--{k}", nargs="+", default=v, type=v_type)
        else:
            parser.add_argument(f"--{k}", default=v, type=v_type)


def args_to_dict(args, keys):
    return {k: getattr(args, k) for k in keys}


def str2bool(v):
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("boolean value expected")
