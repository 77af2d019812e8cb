"""Model selection exactly as the reference's entry points do it."""
from image_denoising_b200 import RESNET, ImprovedUNet, UNet


def network_from_log_name(log_name: str, n_channel: int, n_feature: int):
    """train.py:298-314 / evaluation.py:32-48 / evaluation_704.py:28-44: the family is a substring of --log_name
    (upper-case 'UNET' selects UNet; the reference leaves `network` undefined when nothing matches)."""
    if 'UNET' in log_name and 'blindspot' in log_name:
        return UNet(in_nc=n_channel, out_nc=n_channel, n_feature=n_feature, blindspot=True)     # raises: out of scope
    if 'UNET' in log_name:
        return UNet(in_nc=n_channel, out_nc=n_channel, n_feature=n_feature)
    if 'RESNET' in log_name:
        return RESNET(in_nc=n_channel, out_nc=n_channel, n_feature=n_feature)
    if 'UNetImproved' in log_name:
        return ImprovedUNet(in_nc=n_channel, out_nc=n_channel, n_feature=n_feature)             # §8f N2 (image_denoising_b200/improved.py)
    raise SystemExit(f"--log_name {log_name!r} selects no network (it must contain 'UNET', 'RESNET' or 'UNetImproved'; the reference "
                     "leaves `network` undefined in this case, train.py:298-314)")


def build_base_model(arch: str, n_channel: int, n_feature: int):
    """finetune.py:189-204 / evaluation_adapter.py:47-56."""
    if arch == 'UNet':
        return UNet(in_nc=n_channel, out_nc=n_channel, n_feature=n_feature)
    if arch == 'RESNET':
        return RESNET(in_nc=n_channel, out_nc=n_channel, n_feature=n_feature)
    if arch == 'UNetImproved':
        return ImprovedUNet(in_nc=n_channel, out_nc=n_channel, n_feature=n_feature)
    raise ValueError(f'Unknown arch: {arch}')


def strip_module_prefix(state):
    """finetune.py:210-212 / evaluation_adapter.py:62-63: checkpoints saved under nn.DataParallel."""
    if any(k.startswith('module.') for k in state.keys()):
        return {k.replace('module.', '', 1): v for k, v in state.items()}
    return state
