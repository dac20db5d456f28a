"""Plain-PyTorch restatement of the reference's photometric loss and densification bookkeeping,
statement for statement, runnable on any device.  TEST / BASELINE INFRASTRUCTURE ONLY (the GPU box has no
/root/reference to import): tests compare the CUDA kernels with it on the GPU, scripts/quick_perf_next_rows.py
times it as the "what the reference executes" baseline.  Nothing under splatfields_b200/ imports this module.

  l1_loss / gaussian / create_window / ssim / _ssim     utils/loss_utils.py:18-19, :33-76
  add_densification_stats / max_radii2D update          scene/gaussian_model.py:427-430, train.py:280-282
  gaussian_dict_static (the GaussianModel getters)      scene/gaussian_model.py:53-86, train.py:42-50, :73
"""
from math import exp

import torch
import torch.nn.functional as F


def l1_loss(network_output, gt):
    return torch.abs((network_output - gt)).mean()


def gaussian(window_size, sigma):
    gauss = torch.Tensor([exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    return gauss / gauss.sum()


def create_window(window_size, channel):
    w1 = gaussian(window_size, 1.5).unsqueeze(1)
    w2 = w1.mm(w1.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous()


def ssim(img1, img2, window_size=11):
    channel = img1.size(-3)
    window = create_window(window_size, channel).to(img1.device).type_as(img1)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def photometric_loss(image, gt_image, lambda_dssim, opacity=None, gt_mask=None, lambda_mask=0.0):
    """train.py:183-184, :189-193."""
    Ll1 = l1_loss(image, gt_image)
    loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - ssim(image, gt_image))
    if opacity is not None:
        opacity_image = torch.clamp(opacity, 0.0, 1.0)
        loss = loss + lambda_mask * F.l1_loss(opacity_image.view(-1), gt_mask.view(-1))
    return loss, Ll1


def add_densification_stats(xyz_gradient_accum, denom, max_radii2D, viewspace_grad, radii):
    """scene/gaussian_model.py:427-430 + train.py:280-282 (in place; visibility_filter = radii > 0)."""
    visibility_filter = radii > 0
    max_radii2D[visibility_filter] = torch.max(max_radii2D[visibility_filter], radii[visibility_filter])
    xyz_gradient_accum[visibility_filter] += torch.norm(viewspace_grad[visibility_filter, :2], dim=-1, keepdim=True)
    denom[visibility_filter] += 1


def gaussian_dict_static(xyz, raw_scaling, raw_rotation, raw_opacity, features_dc, features_rest, scale_offset=None,
                         use_isotropic=False):
    """train.py:42-50 through the getters of scene/gaussian_model.py:64-86 (activations set at :53-61), plus the
    `ret['scales'] + scaling` epilogue of the dynamic branch (train.py:73) when scale_offset is given."""
    scaling = torch.exp(raw_scaling)                                   # :53, :64-68
    if use_isotropic:
        scaling = scaling.repeat(1, 3)
    if scale_offset is not None:
        scaling = scale_offset + scaling                               # train.py:73
    return {
        "means3D": xyz,
        "gaussian_opacity": torch.sigmoid(raw_opacity),                # :58, :83-85
        "gaussian_features": torch.cat((features_dc, features_rest), dim=1),   # :78-82
        "gaussian_scales": scaling,
        "gaussian_rotations": F.normalize(raw_rotation),               # :61, :70-72
    }
