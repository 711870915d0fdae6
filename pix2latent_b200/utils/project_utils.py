"""Writes an experiment's artefacts to disk (reference: pix2latent/utils/project_utils.py:13-48; that file
imports a module that does not exist in the package — ``im_utils`` — so it cannot run as shipped; the
intended behaviour is implemented): progress video, target / weight / best result as JPEGs, the loss
history and the variables as .npy pickles."""
import os.path as osp

import numpy as np

from .image import to_image
from .video import make_video


def save_result(save_dir, fn, collages, target, weight, out, vars, losses, t_outs=None, t_out=None, t_target=None,
                transform=None, metric="vgg"):
    import cv2
    q = [int(cv2.IMWRITE_JPEG_QUALITY), 100]
    last = losses[-1][1]
    key = metric if metric in last else sorted(last.keys())[0]  # the reference hard-codes 'vgg'
    idx = int(np.argmin(last[key]))

    def jpg(name, im):
        cv2.imwrite(osp.join(save_dir, "{}.{}.jpg".format(fn, name)), np.ascontiguousarray(im).astype(np.uint8), q)

    make_video(osp.join(save_dir, "{}.mp4".format(fn)), collages, duration=5)
    jpg("target", to_image(target)[0])
    jpg("weight", to_image(weight)[0])
    if t_out is not None:
        jpg("transform.final", to_image(out)[idx])
        jpg("final", to_image(t_out)[idx])
    else:
        jpg("final", to_image(out)[idx])
    if t_target is not None:
        jpg("transform.target", to_image(t_target)[0])
    np.save(osp.join(save_dir, "{}.loss.npy".format(fn)), np.array(losses, dtype=object), allow_pickle=True)
    if t_outs is not None:
        make_video(osp.join(save_dir, "{}.transform.mp4".format(fn)), t_outs[0], duration=5)
        make_video(osp.join(save_dir, "{}.transform.out.mp4".format(fn)), t_outs[1], duration=5)
    np.save(osp.join(save_dir, "{}.vars.npy".format(fn)), {"vars": vars, "transform": transform}, allow_pickle=True)
