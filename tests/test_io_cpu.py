"""Result I/O and logging formats: vars.npy pickle, image helpers, videos, save_result."""
import os

import numpy as np
import torch


def test_save_variables_roundtrip(tmp_path):
    from pix2latent_b200 import VariableManager, save_variables
    vm = VariableManager(device="cpu")
    vm.register("z", (4,), "input")
    vm.register("target", (3, 2, 2), "output", requires_grad=False, default=torch.zeros(3, 2, 2))
    v = vm.initialize(3)
    p = str(tmp_path / "vars.npy")
    save_variables(p, v)
    back = np.load(p, allow_pickle=True).item()
    assert "opt" not in back and back["num_samples"] == 3
    assert len(back["input"]["z"]["data"]) == 3 and not back["input"]["z"]["data"][0].requires_grad
    assert torch.equal(back["input"]["z"]["data"][1], v.input.z.data[1].detach())


def test_image_helpers():
    from pix2latent_b200.utils import image
    rgb = (np.random.RandomState(0).rand(10, 14, 3) * 255).astype(np.uint8)
    t = image.to_tensor(rgb, device="cpu")
    assert t.shape == (1, 3, 10, 14) and -1 <= t.min() and t.max() <= 1
    back = image.to_image(t, cv2_format=False)[0]
    assert np.abs(back.astype(int) - rgb.astype(int)).max() <= 1
    m = image.to_mask((np.random.RandomState(1).rand(10, 14, 1) > 0.5).astype(np.float64), device="cpu")
    assert m.shape == (1, 1, 10, 14) and set(m.unique().tolist()) <= {0.0, 1.0}
    assert image.center_crop(rgb).shape == (10, 10, 3) and image.center_crop(rgb.transpose(1, 0, 2)).shape == (10, 10, 3)
    assert np.array_equal(image.center_crop(rgb), rgb[:, 2:12])
    assert image.smart_resize(rgb, (5, 7)).shape == (5, 7, 3) and image.smart_resize(rgb, (20, 28)).shape == (20, 28, 3)
    g = image.to_grid(torch.zeros(5, 3, 4, 4))
    assert g.shape[0] == 3 and g.shape[1] > 8  # 3 x 3 grid with padding
    target = np.full((32, 32, 3), 0.5)
    gen = np.full((32, 32, 3), 0.2)
    mask = np.zeros((32, 32, 3)); mask[8:24, 8:24] = 1.0
    out = image.poisson_blend(target, mask, gen)
    assert out.shape == (32, 32, 3) and out.dtype == np.uint8


def test_videos_and_save_result(tmp_path):
    from pix2latent_b200.utils import video
    from pix2latent_b200.utils.project_utils import save_result
    frames = [np.full((16, 16, 3), i / 4.0) for i in range(4)]
    gif = str(tmp_path / "a.gif")
    video.make_gif(gif, [(f * 255).astype(np.uint8) for f in frames], duration=1.0)
    assert os.path.getsize(gif) > 0
    assert video.make_video(str(tmp_path / "a.avi"), frames) is False  # unsupported container: prints, returns False
    video.make_video(str(tmp_path / "a.mp4"), frames, duration=1)
    assert os.path.exists(str(tmp_path / "a.mp4"))
    out = torch.rand(3, 3, 16, 16) * 2 - 1
    losses = [[5, {"loss": np.array([0.3, 0.1, 0.2])}]]
    save_result(str(tmp_path), "run", frames, out[:1], out[:1], out, {"k": 1}, losses)
    for name in ("run.mp4", "run.target.jpg", "run.weight.jpg", "run.final.jpg", "run.loss.npy", "run.vars.npy"):
        assert os.path.exists(str(tmp_path / name)), name
    assert np.load(str(tmp_path / "run.vars.npy"), allow_pickle=True).item()["vars"] == {"k": 1}
