"""Loss factory with the reference's public entry point for the adaptation path.

Mirrors `loss_factory.get_reprojection_loss(name, multiScale=False, logs=False, weights=None, reduced=True)` of the
reference (Losses/loss_factory.py:353-395): it returns `compute_loss(disparities, inputs)` which warps `inputs['right']`
with `disparities[-1]` (clamped bilinear sampler, images / 256) and compares it with `inputs['left']`.  Only the loss the
online-adaptation scripts use, 'mean_SSIM_l1' (0.85 * mean SSIM dissimilarity over 3x3 VALID windows + 0.15 * mean L1,
:128-164), exists on the device (`ms_reproj_loss`, csrc/loss.cu); the other names of the reference's tables are recognised
and rejected with NotImplementedError, unknown names raise the reference's Exception.

Inside the engine the same kernel computes the full-resolution and the per-module losses of every frame; this wrapper is
the stand-alone operator form (e.g. to score a disparity map).
"""
import torch

from madstereo import ops

SUPERVISED_LOSS = ['mean_l1', 'sum_l1', 'mean_l2', 'sum_l2', 'mean_SSIM', 'mean_SSIM_l1', 'ZNCC', 'cos_similarity',
                   'smoothness', 'mean_huber', 'sum_huber']
PIXELWISE_LOSSES = ['l1', 'l2', 'SSIM', 'huber', 'ssim_l1']
ALL_LOSSES = SUPERVISED_LOSS + PIXELWISE_LOSSES


def _as_tensor(x):
    if hasattr(x, 'tensor') and callable(x.tensor):        # Nets.Stereo_net.LayerHandle
        x = x.tensor()
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    return x.to(device='cuda', dtype=torch.float32).contiguous()


def get_reprojection_loss(reconstruction_loss, multiScale=False, logs=False, weights=None, reduced=True):
    if reconstruction_loss not in ALL_LOSSES:
        print('Unrecognized loss function, pick one among: {}'.format(ALL_LOSSES))
        raise Exception('Unknown loss function selected')
    if reconstruction_loss != 'mean_SSIM_l1':
        raise NotImplementedError("only 'mean_SSIM_l1' (the loss of the adaptation path) is implemented on the device")
    if multiScale:
        raise NotImplementedError('multiScale=True is not used by the adaptation path (Stereo_Online_Adaptation.py:70,107)')
    if weights is None:
        weights = [1] * 10

    def compute_loss(disparities, inputs):
        left, right = _as_tensor(inputs['left']), _as_tensor(inputs['right'])
        disp = _as_tensor(disparities[-1])
        if disp.shape[1:3] != left.shape[1:3]:
            raise NotImplementedError('the prediction must already have the resolution of the inputs (scale factor 1 on this path)')
        loss, _ = ops.reprojection_loss(left, right, disp, with_grad=False)
        loss = loss * weights[0]
        return loss.reshape(()) if reduced else [loss.reshape(())]
    return compute_loss
