"""ucdir_b200: B200-native (sm_100a) implementation of UCDIR's iterative-denoising hot path.

Public surface = the reference's: `ucdir_b200.model.networks.define_G(opt)` returning a module with
`.super_resolution(x)`; everything below it runs in libucdir_b200.so (hand-written CUDA behind a C ABI,
include/ucdir_b200.h).
"""
__version__ = "0.1.0"

SID_MODEL_OPT = {
    "which_model_G": "ucdir", "unet_name": "DY3h", "diffusion_name": "ResiGaussianGuideDY", "finetune_norm": False,
    "unet": {"in_channel": 6, "out_channel": 3, "inner_channel": 64, "channel_mults": [1, 2, 4, 8, 8],
             "attn_res": [16], "res_blocks": 2, "dropout": 0.1, "norm_groups": 1},
    "beta_schedule": {"train": {"schedule": "linear", "n_timestep": 2000, "linear_start": 1e-6, "linear_end": 0.01},
                      "val": {"schedule": "linear", "n_timestep": 200, "linear_start": 1e-6, "linear_end": 0.1}},
    "diffusion": {"image_size": 128, "channels": 3, "conditional": True},
}
"""config/sid.yaml `model:` section of the reference, restated as a dict (the GPU box has no reference tree)."""

SID_VAL_SCHEDULE = {"schedule": "linear", "n_timestep": 50, "linear_start": 1e-6, "linear_end": 0.4}
"""core/logger.py:58-61 override applied by `sr.py -p val` when 'sid' is in the experiment name."""
