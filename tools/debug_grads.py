"""Debug helper (GPU box): per-parameter gradient error of the engine vs the CPU oracle for a golden fixture."""
import os, sys, random
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "stylegan.pytorch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import stylegan_oracle as O
from test_oracle_golden import arch_from
from test_gpu_networks import build, set_noise, FixedMixing, rel
from models import Losses

name = sys.argv[1]
fx = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)
arch = arch_from(fx["arch"]); depth, alpha = fx["depth"], fx["alpha"]
# oracle
g = {k: v.clone() for k, v in fx["g_init"].items()}; d = {k: v.clone() for k, v in fx["d_init"].items()}
gn, dn = O.param_names(g), O.param_names(d)
for k in gn: g[k].requires_grad_(True)
for k in dn: d[k].requires_grad_(True)
fake_o = O.generator_forward(g, fx["z"], depth, alpha, arch, z2=fx["mix"][0], mixing_cutoff=fx["mix"][1], noises=fx["noises"])
dl_o = O.d_loss_logistic(d, fx["real"], fake_o.detach(), depth, alpha, arch)
dg_o = dict(zip(dn, torch.autograd.grad(dl_o, [d[k] for k in dn], allow_unused=True)))
dl_nor1 = O.d_loss_logistic(d, fx["real"], fake_o.detach(), depth, alpha, arch, r1_gamma=0.0)
dg_o_nor1 = dict(zip(dn, torch.autograd.grad(dl_nor1, [d[k] for k in dn], allow_unused=True)))
gl_o = O.g_loss_logistic(d, fake_o, depth, alpha, arch)
gg_o = dict(zip(gn, torch.autograd.grad(gl_o, [g[k] for k in gn], allow_unused=True)))
# engine
G, D = build(arch, fx["g_init"], fx["d_init"]); G.train(); D.train(); set_noise(G, fx["noises"])
with FixedMixing(fx["mix"][0], fx["mix"][1], depth):
    fake = G(fx["z"].cuda(), depth, alpha)
print("fake rel", rel(fake, fake_o))
L = Losses.LogisticGAN(D)
real = fx["real"].cuda()
dl = L.dis_loss(real, fake_o.cuda(), depth, alpha)
print("d_loss", float(dl), float(dl_o))
dg = torch.autograd.grad(dl, list(D.parameters()), allow_unused=True)
print("--- D grads (with R1) / (without R1)")
dl2 = L.dis_loss(real, fake_o.cuda(), depth, alpha, r1_gamma=0.0)
dg2 = torch.autograd.grad(dl2, list(D.parameters()), allow_unused=True)
for (n, _), a, a2 in zip(D.named_parameters(), dg, dg2):
    if a is None: continue
    print("%-40s %.2e   %.2e" % (n, rel(a, dg_o[n]), rel(a2, dg_o_nor1[n])))
gl = L.gen_loss(None, fake, depth, alpha)
print("g_loss", float(gl), float(gl_o))
gg = torch.autograd.grad(gl, list(G.parameters()), allow_unused=True)
print("--- G grads")
for (n, _), a in zip(G.named_parameters(), gg):
    if a is None: continue
    print("%-55s %.2e" % (n, rel(a, gg_o[n])))
