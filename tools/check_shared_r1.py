"""GPU self-check of the shared real pass (Losses.LogisticGAN.shared_real_pass): the discriminator loss and every
parameter gradient with one shared D(real) graph against the reference's two-pass formulation, fp32 and bf16, with and
without fade-in.  Prints the largest relative difference per case; exit code 0 when all are within tolerance."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stylegan.pytorch_b200"))
t0 = time.time()
import torch
import models
from models.GAN import Discriminator
from models.Losses import LogisticGAN

dev = torch.device("cuda", 0)
ok = True
for prec, tol in (("fp32", 2e-4), ("bf16", 5e-2)):
    models.set_precision(prec)
    for res, depth, alpha, fmap_base, fmap_max in ((16, 2, 0.5, 128, 32), (16, 0, 1.0, 128, 32), (128, 5, 0.3, 512, 32)):
        torch.manual_seed(res + depth)
        d = Discriminator(resolution=res, fmap_base=fmap_base, fmap_max=fmap_max, blur_filter=[1, 2, 1]).to(dev)
        with torch.no_grad():
            for n, p in d.named_parameters():
                if n.endswith(".bias"):
                    p.copy_(0.1 * torch.randn_like(p))
        d.train()
        L = LogisticGAN(d)
        r = 2 ** (depth + 2)
        real, fake = torch.randn(4, 3, r, r, device=dev), torch.randn(4, 3, r, r, device=dev)
        params = [p for p in d.parameters()]
        out = []
        for shared in (False, True):
            models.set_shared_r1(shared)
            loss = L.dis_loss(real, fake, depth, alpha)
            grads = torch.autograd.grad(loss, params, allow_unused=True)
            out.append((float(loss), grads))
        models.set_shared_r1(True)
        dl = abs(out[0][0] - out[1][0]) / abs(out[0][0])
        worst = 0.0
        for a, b in zip(out[0][1], out[1][1]):
            assert (a is None) == (b is None)
            if a is not None:
                worst = max(worst, float((a - b).norm() / (a.norm() + 1e-20)))
        good = dl < tol and worst < tol
        ok &= good
        print("%s res %3d depth %d alpha %.1f: loss %.6f vs %.6f (rel %.1e), worst gradient rel diff %.1e  %s"
              % (prec, res, depth, alpha, out[0][0], out[1][0], dl, worst, "ok" if good else "MISMATCH"), flush=True)
print("elapsed %.1fs" % (time.time() - t0), flush=True)
sys.exit(0 if ok else 1)
