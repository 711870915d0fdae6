"""Image I/O helpers the examples use (reference: pix2latent/utils/image.py). Host-only; not on
the hot path (out of scope for acceleration, kept so the scripts run)."""
import numpy as np
import torch


def _pil():
    from PIL import Image
    return Image


def read(im_path, as_transformed_tensor=False, im_size=512, transform_style=None):
    """Load an RGB image; with ``as_transformed_tensor`` return a [3, im_size, im_size] tensor in
    [-1, 1]: resize+centre-crop (None / 'biggan') or pad-to-square+resize ('stylegan')."""
    from torchvision import transforms
    Image = _pil()
    im = np.array(Image.open(im_path).convert("RGB"))
    h, w = im.shape[:2]
    if np.max(im) <= 1. + 1e-6:
        im = (im * 255).astype(np.uint8)
    im = Image.fromarray(im)
    if not as_transformed_tensor:
        raise ValueError("read(): only as_transformed_tensor=True is supported (as in the reference, "
                         "which fails with an unbound `transform` otherwise)")
    norm = [transforms.ToTensor(), transforms.Normalize([0.5] * 3, [0.5] * 3)]
    if transform_style in (None, "biggan"):
        tf = [transforms.Resize(im_size), transforms.CenterCrop(im_size)] + norm
    elif transform_style in ("stylegan", "stylegan2"):
        if h < w:
            top = (w - h) // 2
            pad = (0, top, 0, w - h - top)
        else:
            left = (h - w) // 2
            pad = (left, 0, h - w - left, 0)
        tf = [transforms.Pad(pad), transforms.Resize(im_size)] + norm
    else:
        raise ValueError("unknown transformation style {}".format(transform_style))
    return transforms.Compose(tf)(im)


def to_grid(x):
    import torchvision
    n = int(np.ceil(np.sqrt(x.size(0))))
    return torchvision.utils.make_grid(x, n, pad_value=-1)


def to_image(output, to_cpu=True, denormalize=True, jpg_format=True, to_numpy=True, cv2_format=True):
    """BCHW (or CHW) tensor in [-1,1] -> BHWC (or HWC) uint-valued array."""
    batched = output.dim() == 4
    t = (output if batched else output.unsqueeze(0)).detach().float()
    if to_cpu:
        t = t.cpu()
    t = t.permute(0, 2, 3, 1)
    if denormalize:
        t = (t + 1.0) / 2.0
    if jpg_format:
        t = (t * 255).int()
    if cv2_format and output.size(-3) > 1:
        t = t[:, :, :, [2, 1, 0]]
    if to_numpy:
        t = t.numpy()
    return t if batched else t.squeeze(0)


def save(save_path, im):
    import cv2
    if isinstance(im, torch.Tensor):
        im = to_image(im, cv2_format=False)
    return cv2.imwrite(save_path, np.ascontiguousarray(im[:, :, [2, 1, 0]]).astype(np.uint8),
                       [int(cv2.IMWRITE_JPEG_QUALITY), 100])


def binarize(mask, min=0.0, max=1.0, eps=1e-3):
    """Continuous mask -> {min, max} by thresholding at 1 - eps."""
    if isinstance(mask, torch.Tensor):
        assert mask.max() <= 1 + 1e-6, mask.max()
        assert mask.min() >= -1 - 1e-6, mask.min()
        return (mask > 1.0 - eps).float().clamp_(min, max)
    if isinstance(mask, np.ndarray):
        m = (mask > 1.0 - eps).astype(float)
        return np.clip(m, min, max, out=m)
    return False


def _default_device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def to_tensor(im, device=None):
    """HWC image (array in [0,1] or [0,255], or a path) -> [1,3,H,W] tensor in [-1,1] on ``device``
    (the reference hard-codes .cuda(), utils/image.py:112-118)."""
    if isinstance(im, str):
        import cv2
        im = cv2.imread(im)[:, :, [2, 1, 0]]
    im = np.asarray(im)
    if np.max(im) > 1:
        im = im / 255.
    t = (2.0 * (torch.from_numpy(np.ascontiguousarray(im)).float() - 0.5)).permute(2, 0, 1)
    return t.unsqueeze(0).to(device or _default_device())


def to_mask(mask, device=None):
    """HW1 mask (array in [0,1], or a path thresholded at 0.5) -> [1,1,H,W] tensor in [0,1]."""
    if isinstance(mask, str):
        import os
        import cv2
        assert os.path.exists(mask)
        mask = (cv2.imread(mask)[:, :, :1] / 255. > 0.5).astype(np.float64)
    mask = np.asarray(mask)
    assert np.max(mask) <= 1.0 and np.min(mask) >= 0.0
    t = torch.from_numpy(np.ascontiguousarray(mask)).permute(2, 0, 1)
    return torch.clamp(t.unsqueeze(0).to(device or _default_device()).float(), 0.0, 1.0)


def center_crop(image):
    """Square centre crop of an HWC array along its longer side."""
    h, w = image.shape[:2]
    side = min(h, w)
    top, left = (h - side) // 2, (w - side) // 2
    out = image[top:top + side, left:left + side, :]
    assert out.shape[0] == out.shape[1]
    return out


def smart_resize(im, target_size=(256, 256)):
    """cv2 resize to (H, W): area interpolation when shrinking, bilinear when enlarging."""
    import cv2
    shrinking = np.prod(im.shape[:2]) >= np.prod(target_size)
    return cv2.resize(im, (target_size[1], target_size[0]), interpolation=cv2.INTER_AREA if shrinking else cv2.INTER_LINEAR)


def poisson_blend(target, mask, generated):
    """Seamless-clone ``generated`` into ``target`` inside ``mask`` (HWC arrays), centred on the mask's box."""
    import cv2
    from ..transform.transform_utils import compute_stat_from_mask
    if np.max(target) <= 1.0:
        target = target * 255.
    if np.max(generated) <= 1.0:
        generated = generated * 255.
    if np.max(mask) > 1.0:
        mask = mask / 255.
    (cy, cx), _ = compute_stat_from_mask(binarize(torch.Tensor(mask).permute(2, 0, 1)))
    center = (int(cx * target.shape[1]), int(cy * target.shape[0]))
    hard = (mask > 0.5).astype(np.float64)
    return cv2.seamlessClone(generated.astype(np.uint8), target.astype(np.uint8), (255 * hard[:, :, 0]).astype(np.uint8),
                             center, cv2.NORMAL_CLONE)
