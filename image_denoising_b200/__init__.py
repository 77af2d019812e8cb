"""image_denoising_b200 — B200-native (sm_100a) drop-in for the Neighbor2Neighbor hot path of
lmh9507/image_denoising: ``arch_unet.UNet``, ``adapter.DenoiserWithAdapter``,
``generate_mask_pair`` / ``generate_subimages``, the N2N / finetune losses, Adam, and the
PSNR/SSIM evaluation, all executed by hand-written CUDA kernels behind the C-ABI of
include/n2n_b200.h.  No CPU or PyTorch-arithmetic fallback exists: importing is cheap, but
every compute entry point needs libn2n_b200.so and a CUDA device."""
from . import _ext  # noqa: F401
from . import dp  # noqa: F401
from .arch_unet import RESNET, ImprovedUNet, UNet  # noqa: F401
from .adapter import DenoiserWithAdapter, OutputAdapter  # noqa: F401
from .n2n import (AugmentNoise, checkpoint, forward_pair, generate_mask_pair, generate_packed_selector,  # noqa: F401
                  generate_subimage_pair, generate_subimages, get_generator, space_to_depth)
from .losses import Structure_loss, iqsl_loss, l1_grad_loss, n2n_loss  # noqa: F401
from .optim import FusedAdam, multistep_lr  # noqa: F401
from .trainer import N2NTrainer  # noqa: F401

__version__ = "0.1.0"
