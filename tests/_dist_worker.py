"""Worker for tests/test_parallel_cpu.py: world_size-2 gloo run of BasinCMAOptimizer with the
oracle model on CPU; rank 0 writes the results."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "golden"))


def fused_toy(out_path, fused):
    """BasinCMA on the toy problem of tests/test_fused_host_cpu.py with the native calls replaced by their CPU
    stand-ins (device-resident loop -> oracle/inner_loop.py, per-step native path -> the autograd path)."""
    sys.path.insert(0, HERE)
    import test_fused_host_cpu as T
    from _pytest.monkeypatch import MonkeyPatch
    from pix2latent_b200.optimizer import BasinCMAOptimizer, closure
    from pix2latent_b200.utils import function_hooks as hk
    mp = MonkeyPatch()
    model, loss_fn = T.ToyModel(), T.ToyLoss()
    T._patch(mp, model, loss_fn)
    mp.setattr(closure, "_step_native", closure._step_autograd)
    vm = T._make(0, hk.Clamp(0.6))
    vm.edit_variable("z", {"grad_free": True})
    opt = BasinCMAOptimizer(model, vm, loss_fn, max_batch_size=3)
    opt.cma_seed = 5
    opt.fuse_inner_loop = fused
    variables, outs, loss = opt.optimize(meta_steps=2, grad_steps=3, last_grad_steps=4)
    res = dict(loss=np.array(loss[0][1]["loss"], dtype=np.float64), z=torch.stack(variables.input.z.data).detach().numpy(),
               c=torch.stack(variables.input.c.data).detach().numpy(), fused_calls=np.array(opt.fused_calls),
               n=np.array(opt.num_samples), tracked=torch.stack(opt.tracked["z"]).numpy(),
               mean=np.array(list(opt.cma_optimizers.values())[0].mean()))
    mp.undo()
    if not dist.is_initialized() or dist.get_rank() == 0:
        np.savez(out_path, **res)
    return res


def transform_search(out_path):
    """TransformBasinCMAOptimizer on the tiny problem of tests/golden/make_golden_transform.py with the torch resampler."""
    import make_golden as mg
    import make_golden_transform as mgt
    from oracle import lpips as olp
    from oracle import transform as otf
    from pix2latent_b200 import VariableManager
    from pix2latent_b200.transform import TransformBasinCMAOptimizer
    import pix2latent_b200.distribution as dst
    import pix2latent_b200.utils.function_hooks as hook
    cfg, model, target, weight = mg.problem()
    loss_fn = olp.ProjectionLoss(lpips_module=olp.make_lpips("alex", seed=0))
    torch.manual_seed(31)
    vm = VariableManager(device="cpu")
    mgt.register_transform_problem(vm, hook, dst, model, target, weight)
    opt = TransformBasinCMAOptimizer(model, vm, loss_fn, max_batch_size=4)
    opt.cma_seed = mg.CMA_SEED
    opt.register_transform(otf.TorchSpatialTransform(t=[1.0, 0.0, 0.0]), "t", "target")
    opt.register_transform(otf.TorchSpatialTransform(t=[1.0, 0.0, 0.0]), "t", "weight")
    opt.set_variable_propagation("z")
    variables, (t_out, t_target, t_cand), loss = opt.optimize(meta_steps=3, grad_steps=2)
    if dist.get_rank() == 0:
        np.savez(out_path, loss=np.array(loss, dtype=np.float64), z=torch.stack(variables.input.z.data).detach().numpy(),
                 tracked=torch.stack(opt.transform_tracked).numpy(), cand=opt.get_candidate().numpy(),
                 vp=opt.vp_means["z"].numpy())


def native_basincma(out_path):
    """BasinCMAOptimizer with the NATIVE BigGAN / loss (tiny128 config) — single process, or one rank per GPU under NCCL
    (LOCAL_RANK picks the device). Every rank evaluates its shard; rank 0 writes final losses / latents."""
    import make_golden as mg
    import test_step_gpu as ts
    from oracle import lpips as olp
    from pix2latent_b200 import VariableManager
    from pix2latent_b200.loss_functions import ProjectionLoss
    from pix2latent_b200.model import BigGAN
    from pix2latent_b200.optimizer import BasinCMAOptimizer
    import pix2latent_b200.distribution as dst
    import pix2latent_b200.utils.function_hooks as hook
    ts._setup()
    cfg, orc, target, weight = mg.problem()
    lp = olp.make_lpips("alex", seed=0)
    model = BigGAN(config=ts._product_cfg(cfg), state_dict=orc.state_dict())
    loss_fn = ProjectionLoss(lpips_state_dict={k: v.cuda() for k, v in ts._lpips_state(lp).items()})
    torch.manual_seed(23)
    vm = VariableManager(device="cuda")
    mg.register(vm, hook, dst, model, target.cuda(), weight.cuda(), True)
    opt = BasinCMAOptimizer(model, vm, loss_fn, max_batch_size=9)
    opt.cma_seed = mg.CMA_SEED
    trace = {"asked": [], "told": []}   # per meta-iteration: what CMA asked, which losses it was told
    init0, upd0 = opt.cma_init, opt.cma_update

    def init1(*a, **k):
        v = init0(*a, **k)
        trace["asked"].append(np.array(list(opt._sampled.values())[0], dtype=np.float64))
        return v

    def upd1(*a, **k):
        l = upd0(*a, **k)
        trace["told"].append(np.array(l, dtype=np.float64))
        return l

    opt.cma_init, opt.cma_update = init1, upd1
    variables, outs, loss = opt.optimize(meta_steps=2, grad_steps=3, last_grad_steps=4)
    if not dist.is_initialized() or dist.get_rank() == 0:
        np.savez(out_path, loss=np.array(loss[0][1]["loss"], dtype=np.float64), z=torch.stack(variables.input.z.data).detach().cpu().numpy(),
                 asked=np.stack(trace["asked"]), told=np.stack(trace["told"]),
                 c=torch.stack(variables.input.c.data).detach().cpu().numpy(), mean=np.array(list(opt.cma_optimizers.values())[0].mean()),
                 fused_calls=np.array(opt.fused_calls), world=np.array(dist.get_world_size() if dist.is_initialized() else 1))


def main():
    out_path = sys.argv[1]
    if len(sys.argv) > 2 and sys.argv[2] == "native":
        torch.set_num_threads(4)
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
        if int(os.environ.get("WORLD_SIZE", 1)) > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
        native_basincma(out_path)
        if dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
        return
    dist.init_process_group("gloo")
    torch.set_num_threads(4)
    if len(sys.argv) > 2 and sys.argv[2] == "transform":
        transform_search(out_path)
        dist.barrier()
        dist.destroy_process_group()
        return
    if len(sys.argv) > 2 and sys.argv[2] == "fused":
        fused_toy(out_path, True)
        dist.barrier()
        dist.destroy_process_group()
        return
    import make_golden as mg
    from oracle import lpips as olp
    from pix2latent_b200 import VariableManager
    from pix2latent_b200.optimizer import BasinCMAOptimizer
    import pix2latent_b200.distribution as dst
    import pix2latent_b200.utils.function_hooks as hook
    cfg, model, target, weight = mg.problem()
    loss_fn = olp.ProjectionLoss(lpips_module=olp.make_lpips("alex", seed=0))
    torch.manual_seed(23)
    vm = VariableManager(device="cpu")
    mg.register(vm, hook, dst, model, target, weight, True)
    opt = BasinCMAOptimizer(model, vm, loss_fn, max_batch_size=9)
    opt.cma_seed = mg.CMA_SEED
    variables, outs, loss = opt.optimize(meta_steps=2, grad_steps=2, last_grad_steps=2)
    if dist.get_rank() == 0:
        np.savez(out_path, loss=np.array(loss[0][1]["loss"]), z=torch.stack(variables.input.z.data).detach().numpy(),
                 mean=np.array(list(opt.cma_optimizers.values())[0].mean()), grid=np.array(outs[0].shape))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
