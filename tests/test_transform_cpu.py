"""Transformation search on CPU, against vectors from the REAL reference code
(tests/golden/make_golden_transform.py -> reference_transform_cpu.npz):
  (a) the oracle's explicit-arithmetic resampler (same formulas as the CUDA kernel) reproduces the
      reference SpatialTransform (forward, inverse, round trip, out-of-image samples);
  (b) the product's pre-alignment arithmetic reproduces compute_pre_alignment / bbox_from_mask;
  (c) the product's TransformBasinCMAOptimizer (host loop, apply_transform, variable propagation, inverted-loss
      tell) reproduces the reference's run when driven with the same model / loss / transform callables."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_golden as mg  # noqa: E402
import make_golden_transform as mgt  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "reference_transform_cpu.npz"))


def test_oracle_resampler_matches_reference():
    from oracle import transform as otf
    ims, delta, _ = mgt.transform_inputs()
    st = otf.SpatialTransform(t=[1.1, 0.05, -0.1], sensitivity=0.1)
    np.testing.assert_allclose(st(ims, delta).numpy(), GOLD["st_fwd"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(st(ims, delta, invert=True).numpy(), GOLD["st_inv"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(st(st(ims, delta), delta, invert=True).numpy(), GOLD["st_roundtrip"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(st(ims[:1], delta).numpy(), GOLD["st_shared_src"], rtol=1e-5, atol=1e-5)
    # identity row reproduces the image; the far-shifted row is (almost) all padding
    np.testing.assert_allclose(GOLD["st_fwd"][0], otf.affine_resample(ims[:1], otf.theta_of(torch.tensor([[1.1, 0.05, -0.1]]))).numpy()[0],
                               rtol=1e-5, atol=1e-5)
    assert np.mean(GOLD["st_fwd"][3] == 0.0) > 0.5


def test_product_theta_matches_oracle():
    from oracle import transform as otf
    from pix2latent_b200.transform.spatial_transform import _theta
    t = torch.tensor([[1.2, 0.3, -0.2], [0.7, -0.5, 0.1]])
    for inv in (False, True):
        assert torch.equal(_theta(t, inv), otf.theta_of(t, inv))


def test_product_prealignment_matches_reference():
    from pix2latent_b200.transform import SpatialTransform
    from pix2latent_b200.transform.transform_utils import bbox_from_mask, compute_pre_alignment
    _, _, mask = mgt.transform_inputs()
    np.testing.assert_array_equal(np.asarray(bbox_from_mask(mask)), GOLD["bbox"])
    np.testing.assert_allclose(compute_pre_alignment(mask), GOLD["prealign_t"], rtol=1e-6)
    np.testing.assert_allclose(SpatialTransform(pre_align=mask).get_default_param().numpy(), GOLD["prealign_default"], rtol=1e-6)
    assert bbox_from_mask(torch.zeros(3, 8, 8)) == (0, 0, 8, 8)  # empty mask: whole range


def test_spatial_transform_has_no_cpu_path():
    from pix2latent_b200.transform import SpatialTransform
    with pytest.raises(RuntimeError, match="CUDA"):
        SpatialTransform()(torch.zeros(1, 3, 8, 8), torch.zeros(1, 3))


def test_product_transform_optimizer_matches_reference():
    from oracle import lpips as olp
    from oracle import transform as otf
    from pix2latent_b200 import VariableManager
    from pix2latent_b200.transform import TransformBasinCMAOptimizer
    import pix2latent_b200.distribution as dist
    import pix2latent_b200.utils.function_hooks as hook
    cfg, model, target, weight = mg.problem()
    loss_fn = olp.ProjectionLoss(lpips_module=olp.make_lpips("alex", seed=0))
    torch.manual_seed(31)
    vm = VariableManager(device="cpu")
    mgt.register_transform_problem(vm, hook, dist, model, target, weight)
    opt = TransformBasinCMAOptimizer(model, vm, loss_fn, max_batch_size=4)
    opt.cma_seed = mg.CMA_SEED
    opt.register_transform(otf.TorchSpatialTransform(t=[1.0, 0.0, 0.0]), "t", "target")
    opt.register_transform(otf.TorchSpatialTransform(t=[1.0, 0.0, 0.0]), "t", "weight")
    opt.set_variable_propagation("z")
    variables, (t_out, t_target, t_candidate), loss = opt.optimize(meta_steps=3, grad_steps=2)
    assert opt.num_samples == int(GOLD["tb_num_samples"]) == 7
    assert len(opt.tracked["z"]) == int(GOLD["tb_tracked_len"])
    np.testing.assert_allclose(torch.stack(opt.transform_tracked).numpy(), GOLD["tb_transform_tracked"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.array(loss, dtype=np.float64), GOLD["tb_loss"], rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(opt.get_candidate().numpy(), GOLD["tb_candidate_t"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(float(opt._best_loss), float(GOLD["tb_best_loss"]), rtol=2e-4)
    np.testing.assert_allclose(torch.stack(variables.input.z.data).detach().numpy(), GOLD["tb_z"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(opt.vp_means["z"].numpy(), GOLD["tb_vp_mean_z"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(float(t_candidate.mean()), float(GOLD["tb_candidate_target_mean"]), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(np.array(list(opt.cma_optimizers.values())[0].mean()), GOLD["tb_cma_mean"], rtol=1e-3, atol=1e-4)
    assert tuple(t_target[0].shape) == tuple(GOLD["tb_target_grid_shape"])
    # del_variable_propagation removes a propagated variable (the reference's version cannot, see its source)
    opt.del_variable_propagation("z")
    assert opt.variables_to_propagate == []
