"""Diagnostic: how much of native-vs-oracle error is bf16 rounding? Compares against a bf16
autocast run of the oracle itself."""
import sys, torch
sys.path.insert(0, ".")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from oracle.biggan import BigGANConfig, make_biggan
from pix2latent_b200.native import NativeBigGAN

def rel(a, b): return ((a.double()-b.double()).norm()/(b.double().norm()+1e-30)).item()
def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten(); return (a@b/(a.norm()*b.norm())).item()

cfg = BigGANConfig.tiny128() if len(sys.argv) < 2 else BigGANConfig.deep256()
orc = make_biggan(cfg, seed=0).cuda()
nat = NativeBigGAN(cfg, orc.state_dict())
torch.manual_seed(2)
b = 5
z = torch.fmod(torch.randn(b, 128), 2.0).cuda().requires_grad_(True)
c = orc.get_class_embedding(3).repeat(b, 1).clone().requires_grad_(True)
ref = orc(z=z, c=c)
dimg = torch.randn_like(ref) * 1e-3
ref.backward(dimg)
gz, gc = z.grad.clone(), c.grad.clone()
z.grad = None; c.grad = None
with torch.autocast("cuda", dtype=torch.bfloat16):
    ab = orc(z=z, c=c)
ab.float().backward(dimg)
print("autocast-bf16 oracle vs fp32 oracle: img max", (ab.float()-ref).abs().max().item(), "rel", rel(ab, ref),
      "dz rel", rel(z.grad, gz), "cos", cos(z.grad, gz), "dc rel", rel(c.grad, gc))
img = nat.forward(z.detach(), c.detach())
dz, dc = nat.backward(b, dimg)
torch.cuda.synchronize()
print("native vs fp32 oracle:               img max", (img-ref).abs().max().item(), "rel", rel(img, ref),
      "dz rel", rel(dz, gz), "cos", cos(dz, gz), "dc rel", rel(dc, gc), "cos", cos(dc, gc))
print("native vs autocast oracle: img rel", rel(img, ab.float()))
from pix2latent_b200 import _lib
for gs in (1, 64, 1024, 4096, 65536, 1 << 20, 1 << 24):
    _lib.set_option("grad_scale", gs)
    dz, dc = nat.backward(b, dimg)
    torch.cuda.synchronize()
    print("grad_scale %8d: dz rel %.4f cos %.5f  dc rel %.4f" % (gs, rel(dz, gz), cos(dz, gz), rel(dc, gc)))
