"""Result videos of the examples (reference: pix2latent/utils/video.py:14-70): a list of HWC frames ->
.gif / .webm / .mp4. Host-only. The reference uses imageio (gif) and scikit-video (mp4); neither is a hard
dependency here: gif goes through PIL when imageio is absent, mp4 through OpenCV's writer when scikit-video
is absent."""
import numpy as np


def _frames_u8(ims):
    ims = np.array(ims)
    if np.max(ims) <= 1:
        ims = ims * 255
    return ims.astype(np.uint8)


def make_gif(save_path, ims, duration=20.0):
    """dump a list of images into a gif of ``duration`` seconds"""
    dpf = duration / len(ims)
    try:
        import imageio
        imageio.mimsave(save_path, ims, duration=dpf)
    except ImportError:
        from PIL import Image
        frames = [Image.fromarray(f) for f in _frames_u8(ims)]
        frames[0].save(save_path, save_all=True, append_images=frames[1:], duration=int(1000 * dpf), loop=0)


def make_video(save_path, ims, fps=30, duration=None, safe=True):
    """Video from an array of RGB frames; ``duration`` (seconds), when given, overrides ``fps``.
    Returns False (after printing) for an unsupported container, like the reference."""
    ims = _frames_u8(ims)
    if duration is not None:
        fps = len(ims) / duration
    height, width = ims[0].shape[:2]
    import cv2
    if save_path.endswith("webm"):
        codec = "VP90"
    elif save_path.endswith("mp4"):
        try:
            import skvideo.io
            skvideo.io.vwrite(save_path, ims, inputdict={"-r": str(fps)},
                              outputdict={"-r": str(fps), "-pix_fmt": "yuv420p", "-b": "40000000"})
            print("saved video to {}".format(save_path))
            return
        except ImportError:
            codec = "mp4v"
    else:
        print("unsupported video format")
        return False
    writer = cv2.VideoWriter(save_path, cv2.VideoWriter_fourcc(*codec), fps, (width, height))
    for im in ims:
        writer.write(np.ascontiguousarray(im[:, :, [2, 1, 0]]))
    writer.release()
    print("saved video to {}".format(save_path))
