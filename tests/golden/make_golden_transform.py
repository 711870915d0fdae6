"""Golden vectors for the transformation search, produced by the REAL reference code
(/root/reference/pix2latent/transform/*) on CPU — same stubbing as make_golden.py (missing third-party
modules replaced, ``.cuda()`` neutralised), reference package imported UNMODIFIED.

Pinned byte-for-byte reference code:
  * ``SpatialTransform.transform`` / ``invert_transform`` / ``__call__`` (spatial_transform.py:43-104: affine_grid
    + grid_sample with torch's defaults) on random images and parameters;
  * ``compute_pre_alignment`` / ``bbox_from_mask`` / ``convert_to_t`` (transform_utils.py:53-119);
  * ``TransformBasinCMAOptimizer.optimize`` (transform_optimizer.py:165-255) incl. variable propagation,
    apply_transform on the first inner step, and the inverted-loss ``tell`` (base_cma_optimizer.py:115-140),
    driven with the oracle generator (tiny128) and the reference's own ProjectionLoss.

Run here (CPU container):  python tests/golden/make_golden_transform.py  -> tests/golden/reference_transform_cpu.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

REF = mg.REF


def transform_inputs():
    """Images / parameters shared with the tests."""
    g = torch.Generator().manual_seed(5)
    ims = torch.tanh(torch.randn(4, 3, 24, 20, generator=g))
    delta = torch.randn(4, 3, generator=g)
    delta[0] = 0.0  # identity row
    delta[3] = torch.tensor([6.0, 9.0, -7.0])  # large shift: samples leave the image (zero padding)
    mask = torch.zeros(3, 32, 32)
    mask[:, 7:23, 10:30] = 1.0
    return ims, delta, mask


def register_transform_problem(vm, hook, dist, model, target, weight):
    mg.register(vm, hook, dist, model, target, weight, grad_free=False)
    vm.register(variable_name="t", shape=(3,), requires_grad=False, var_type="transform", grad_free=True)


def main():
    mg.install_stubs()
    sys.path.insert(0, REF)
    import pix2latent
    assert pix2latent.__file__.startswith(REF)
    from pix2latent import VariableManager
    from pix2latent.transform import SpatialTransform, TransformBasinCMAOptimizer
    from pix2latent.transform.transform_utils import compute_pre_alignment, bbox_from_mask
    import pix2latent.loss_functions as LF
    import pix2latent.utils.function_hooks as hook
    import pix2latent.distribution as dist

    out = {}
    ims, delta, mask = transform_inputs()
    st = SpatialTransform(t=[1.1, 0.05, -0.1], sensitivity=0.1)
    out["st_fwd"] = st(ims, delta).numpy()
    out["st_inv"] = st(ims, delta, invert=True).numpy()
    out["st_roundtrip"] = st(st(ims, delta), delta, invert=True).numpy()
    out["st_shared_src"] = st(ims[:1].repeat(4, 1, 1, 1), delta).numpy()
    out["prealign_t"] = np.asarray(compute_pre_alignment(mask), dtype=np.float32)
    out["bbox"] = np.asarray(bbox_from_mask(mask), dtype=np.int64)
    st2 = SpatialTransform(pre_align=mask)
    out["prealign_default"] = st2.get_default_param(as_tensor=True).numpy()

    # ---- TransformBasinCMAOptimizer on the tiny problem
    cfg, model, target, weight = mg.problem()
    loss_fn = LF.ProjectionLoss()
    torch.manual_seed(31)
    vm = VariableManager()
    register_transform_problem(vm, hook, dist, model, target, weight)
    opt = TransformBasinCMAOptimizer(model, vm, loss_fn, max_batch_size=4)
    opt.register_transform(SpatialTransform(t=[1.0, 0.0, 0.0]), "t", "target")
    opt.register_transform(SpatialTransform(t=[1.0, 0.0, 0.0]), "t", "weight")
    opt.set_variable_propagation("z")
    variables, (t_out, t_target, t_candidate), loss = opt.optimize(meta_steps=3, grad_steps=2)
    out["tb_num_samples"] = np.array(opt.num_samples)
    out["tb_loss"] = np.array(loss, dtype=np.float64)
    out["tb_transform_tracked"] = torch.stack(opt.transform_tracked).numpy()
    out["tb_candidate_t"] = opt.get_candidate().numpy()
    out["tb_best_loss"] = np.array(opt._best_loss, dtype=np.float64)
    out["tb_z"] = torch.stack(variables.input.z.data).detach().numpy()
    out["tb_vp_mean_z"] = opt.vp_means["z"].numpy()
    out["tb_candidate_target_mean"] = np.array(t_candidate.mean().item())
    out["tb_target_grid_shape"] = np.array(t_target[0].shape)
    out["tb_cma_mean"] = np.array(list(opt.cma_optimizers.values())[0].mean())
    out["tb_tracked_len"] = np.array(len(opt.tracked["z"]))

    path = os.path.join(HERE, "reference_transform_cpu.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
