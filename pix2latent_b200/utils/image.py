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
