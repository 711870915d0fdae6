"""Metrics evaluated while optimising (reference: pix2latent/utils/benchmark.py:12-46): a dict of loss
objects called as ``metric(out, target, mask)``; perceptual metrics are created on first use."""
import torch


class Benchmark():

    def __init__(self, metrics):
        from .. import loss_functions as LF
        makers = {
            "l1": lambda: LF.ReconstructionLoss(loss_type="l1"),
            "l2": lambda: LF.ReconstructionLoss(loss_type="l2"),
            "alex": lambda: LF.PerceptualLoss("alex"),
            "vgg": lambda: LF.PerceptualLoss("vgg"),
        }
        self._makers, self.metrics = {}, {}
        for m in metrics:
            if m not in makers:
                # (the reference also lists 'squeeze'; the native LPIPS has the alex and vgg backbones)
                raise ValueError("Invalid metric {}".format(m))
            self._makers[m] = makers[m]

    def evaluate(self, out, target, mask):
        result = {}
        with torch.no_grad():
            out = out.cuda()
            for name, make in self._makers.items():
                if name not in self.metrics:
                    self.metrics[name] = make()
                result[name] = self.metrics[name](out, target, mask).detach().cpu().numpy()
        return result
