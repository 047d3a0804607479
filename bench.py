#!/usr/bin/env python
"""Benchmark of the StyleGAN G+D training step (BASELINE.json metric: images/sec @256x256, batch 32 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--res 256] [--batch 32] [--precision bf16]

One "step" = StyleGAN.optimize_discriminator + StyleGAN.optimize_generator (logistic loss with R1, EMA, style mixing,
truncation; reference models/GAN.py:591-659) on synthetic NCHW batches with random-init weights.  For N > 1 launch under
torch.distributed.run (one rank per GPU, NCCL); per-GPU work is fixed (weak scaling), `value` is the whole-job
images/sec from the max-over-ranks device time.  Prints ONE JSON line on rank 0.

`--impl reference` times the CPU oracle port of the reference path (oracle/stylegan_oracle.py; the reference itself
is a Python package that does not exist on the GPU box) on the host cores with a bounded batch per step."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "stylegan.pytorch_b200"))
sys.dont_write_bytecode = True

F_STEP_GFLOP = {128: 656.7, 256: 822.8, 1024: 1162.0}     # BASELINE.md section 3 (reference formulation, per image)
# executed by this engine: 4 F_G + 11 F_D (the logistic term and the R1 penalty share one D(real) graph, Losses.py)
F_EXEC_GFLOP = {128: 547.3, 256: 685.7, 1024: 968.4}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled every 200 ms during the timed region.  In-process NVML (pynvml) on a
    thread: spawning `nvidia-smi -lms` stalled kernel launches for tens of ms at start-up/teardown and perturbed
    the very numbers it annotates; nvidia-smi stays as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.path = None
        self.thread = None
        self.samples = []
        self.reasons = set()
        self._stop = False
        self.period = float(os.environ.get("SG_BENCH_SAMPLE_MS", "200")) * 1e-3

    def _nvml_loop(self, nv, h):
        while not self._stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
            self.sm_max = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            import threading
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            t0 = time.time()
            while time.time() - t0 < 5.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop = True
            self.thread.join(timeout=2)
            sm = sorted(self.samples)
            return {"sm_mhz": float(sm[len(sm) // 2]) if sm else None, "sm_max_mhz": float(self.sm_max),
                    "samples": len(sm), "reasons": sorted(self.reasons), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def make_gan(res, precision, device, seed):
    import torch
    import models
    from models.GAN import StyleGAN
    models.set_precision(precision)
    torch.manual_seed(seed)
    g_args = dict(latent_size=512, mapping_layers=4, blur_filter=[1, 2, 1], truncation_psi=0.7, truncation_cutoff=8)
    d_args = dict(use_wscale=True, blur_filter=[1, 2, 1])
    opt = dict(learning_rate=0.003, beta_1=0, beta_2=0.99, eps=1e-8)           # config.py:79-92
    sg = StyleGAN(structure="linear", resolution=res, num_channels=3, latent_size=512, g_args=g_args, d_args=d_args,
                  g_opt_args=opt, d_opt_args=dict(opt), loss="logistic", use_ema=True, ema_decay=0.999, device=device)
    sg.gen.train(); sg.dis.train(); sg.gen_shadow.train()
    # the reference initialises noise weights to zero (CustomLayers.py:188); randomise so the noise path does work.
    # Drawn from a rank-independent CPU generator so that data-parallel replicas stay identical after the rank-0
    # broadcast StyleGAN.__init__ has already done.
    gen = torch.Generator().manual_seed(4321)
    with torch.no_grad():
        for n, p in sg.gen.named_parameters():
            if n.endswith("noise.weight"):
                p.copy_(torch.randn(p.shape, generator=gen).to(p.device))
        sg.gen_shadow.load_state_dict(sg.gen.state_dict())
    from models import _ops
    _ops.clear_weight_cache()
    return sg


def northstar_kernel_roofline(batch, res, ch, iters=20):
    """The fused 3x3 synthesis conv (blocks[-1].conv1: ch->ch at res x res) timed alone with CUDA events: the CTA-pair
    tcgen05 kernel with bias, noise, LeakyReLU and the InstanceNorm statistics in its epilogue (csrc/sg_conv_pair_tc.cu);
    falls back to the single-CTA kernel (bias + LeakyReLU only) where the pair kernel does not cover the shape."""
    import torch
    from models import _tc
    from models._lib import lib, ptr, stream
    dev = "cuda"
    pair = bool(lib.sg_conv3x3_pair_supported(batch, res, res, ch, ch))
    if not pair and not lib.sg_conv3x3_tc_supported(batch, res, res, ch, ch):
        return None
    x = torch.randn(batch, res, res, ch, device=dev).bfloat16()
    W = torch.nn.Parameter(torch.randn(ch, ch, 3, 3, device=dev))
    b = torch.randn(ch, device=dev)
    y = torch.empty_like(x)
    E = _tc.tc_weight(W, "fwd", 0.05)
    st = stream()
    if pair:
        noise = torch.randn(batch, 1, res, res, device=dev)
        nw = torch.randn(ch, device=dev)
        part = torch.empty(batch, lib.sg_conv3x3_pair_slots(batch, res, res, ch, ch), 2, ch, device=dev)

        def run():
            lib.sg_conv3x3_pair(ptr(x), ptr(E), ptr(y), ptr(b), 1.0, 1, ptr(noise), ptr(nw), ptr(part), batch, res, res,
                                ch, ch, st)
        name = "conv3x3_pair_kernel<stats> %d->%d @%dx%d N=%d (cta_group::2; fused bias + noise + LeakyReLU + " \
               "InstanceNorm partial sums)" % (ch, ch, res, res, batch)
        extra_bytes = noise.numel() * 4 + part.numel() * 4
    else:
        def run():
            lib.sg_conv3x3_tc(ptr(x), ptr(E), ptr(y), ptr(b), 1.0, 1, batch, res, res, ch, ch, st)
        name = "conv3x3_tc_kernel %d->%d @%dx%d N=%d (fused bias+LeakyReLU)" % (ch, ch, res, res, batch)
        extra_bytes = 0
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    # input+output (2 x 268 MB at the north-star shape) exceed the 126 MB L2, so successive launches do not hit in L2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * batch * res * res * ch * ch * 9
    pk, how = peaks()
    ach = flops / (ms * 1e-3) / 1e12
    alg_bytes = 2.0 * batch * res * res * ch * 2 + 9 * ch * ch * 2 + extra_bytes
    return {"bound": "tensor", "kernel": name,
            "achieved": round(ach, 1), "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": round(ach / pk["bf16_tflops"], 4),
            "peak_source": how + " (burst cuBLAS bf16; kernel timed alone)", "ms_per_launch": round(ms, 4),
            "algorithmic_gflop_per_launch": round(flops / 1e9, 1),
            "algorithmic_bytes_per_launch": int(alg_bytes),
            "hbm_gbs_algorithmic": round(alg_bytes / (ms * 1e-3) / 1e9, 1),
            **northstar_traffic(batch, res, ch, "pair" if pair else "conv3x3_tc")}


def northstar_traffic(batch, res, ch, kernel=""):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the north-star kernel from the committed
    `ncu --set full` capture (profiles/northstar_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep);
    null when no capture exists for this shape."""
    try:
        rec = json.load(open(os.path.join(ROOT, "profiles", "northstar_traffic.json")))
        for r in rec["captures"]:
            if (r["batch"], r["res"], r["ch"]) == (batch, res, ch) and kernel in r["kernel"]:
                return {"traffic": r["dram_bytes"], "traffic_unit": "bytes/launch", "traffic_source": r["source"]}
    except Exception:
        pass
    return {"traffic": None, "traffic_unit": "bytes/launch", "traffic_source": None}


def cpu_oracle_step_time(res, batch, steps, warmup, threads=None, budget_s=None):
    """Times the oracle port of the reference's G+D step on the host cores.  -> (seconds/step list, threads).
    budget_s bounds the leg on a slow or shared host (a 256x256 step took 12 s on one box and 60-128 s on another): once
    the elapsed time (warm-up included) exceeds it, the loop stops after the current step -- at least one timed step is
    always taken, and the caller reports how many were."""
    import random
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import stylegan_oracle as O
    # torch.distributed.run exports OMP_NUM_THREADS=1: the CPU arm must still use every host core
    torch.set_num_threads(threads or os.cpu_count() or 1)
    torch.manual_seed(0); random.seed(0)
    arch = O.Arch(res)
    gen = torch.Generator().manual_seed(0)
    st = O.TrainState(arch, O.init_generator(arch, gen), O.init_discriminator(arch, gen))
    depth = arch.depth - 1
    times = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        if budget_s is not None and times and time.perf_counter() - t_begin > budget_s:
            break
        real = torch.randn(batch, 3, res, res)
        z = torch.randn(batch, 512)
        noises_d = [torch.randn(s) for s in O.noise_shapes(arch, batch, depth)]
        noises_g = [torch.randn(s) for s in O.noise_shapes(arch, batch, depth)]
        mix_d, mix_g = O.draw_mixing(arch, z.shape, depth), O.draw_mixing(arch, z.shape, depth)
        t0 = time.perf_counter()
        O.train_step(st, real, z, depth, 0.5, mix_d, mix_g, noises_d, noises_g)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times, torch.get_num_threads()


def gpu_oracle_step_time(res, batch, steps, warmup, mode):
    """The honest GPU comparator: the SAME eager-PyTorch formulation the reference executes (the oracle
    port = the reference's functions on flat state dicts), run on this B200 through cuDNN / cuBLAS instead of our
    kernels.  mode 'fp32': plain fp32 with cudnn.benchmark=True (reference train.py:77); mode 'bf16': the same under
    torch.autocast(bfloat16) with channels_last inputs and weights.  None of this library's kernels run here.
    -> ms/step (CUDA events)."""
    import random
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import stylegan_oracle as O
    dev = torch.device("cuda")
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0); random.seed(0)
    arch = O.Arch(res)
    gen = torch.Generator().manual_seed(0)
    g_sd = {k: v.to(dev) for k, v in O.init_generator(arch, gen).items()}
    d_sd = {k: v.to(dev) for k, v in O.init_discriminator(arch, gen).items()}
    cl = mode == "bf16"
    if cl:
        for sd in (g_sd, d_sd):
            for k, v in sd.items():
                if v.dim() == 4:
                    sd[k] = v.contiguous(memory_format=torch.channels_last)
    st = O.TrainState(arch, g_sd, d_sd)
    depth = arch.depth - 1
    ev = []
    for i in range(warmup + steps):
        real = torch.randn(batch, 3, res, res, device=dev)
        if cl:
            real = real.contiguous(memory_format=torch.channels_last)
        z = torch.randn(batch, 512, device=dev)
        nd = [torch.randn(s, device=dev) for s in O.noise_shapes(arch, batch, depth)]
        ng = [torch.randn(s, device=dev) for s in O.noise_shapes(arch, batch, depth)]
        md, mg = O.draw_mixing(arch, z.shape, depth), O.draw_mixing(arch, z.shape, depth)
        md, mg = (md[0].to(dev), md[1]), (mg[0].to(dev), mg[1])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if cl:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                O.train_step(st, real, z, depth, 0.5, md, mg, nd, ng)
        else:
            O.train_step(st, real, z, depth, 0.5, md, mg, nd, ng)
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            ev.append(e0.elapsed_time(e1))
    del st, g_sd, d_sd
    torch.cuda.empty_cache()
    return sum(ev) / len(ev)


def dp_self_check(sg, device, world, rank, depth, alpha, B, R):
    """Data-parallel invariants, checked on the NCCL path before timing:
      (1) the all-reduced gradients equal the mean of the per-rank gradients (gathered and averaged independently);
      (2) after two optimizer steps every rank holds bit-identical G, D and shadow parameters and avg_latent.
    -> "ok" or a description of the first violation (same string on every rank)."""
    import torch
    import torch.distributed as dist

    def digest(mods):
        acc = []
        for m in mods:
            for t in list(m.parameters()) + list(m.buffers()):
                acc.append(t.detach().double().sum().reshape(1))
                acc.append((t.detach().double() * 1.000001).abs().sum().reshape(1))
        return torch.cat(acc)

    msg = "ok"
    # (1) gradient averaging: run a D backward by hand, keep the local grads, then let the exchange run
    torch.manual_seed(7 + rank)
    img = torch.randn(B, 3, R, R, device=device)
    z = torch.randn(B, 512, device=device)
    with torch.no_grad():
        fake = sg.gen(z, depth, alpha).detach()
        real = sg._StyleGAN__progressive_down_sampling(img, depth, alpha)
    loss = sg.loss.dis_loss(real, fake, depth, alpha)
    sg.dis_optim.zero_grad()
    loss.backward()
    params = [p for p in sg.dis.parameters() if p.grad is not None]
    local = torch.cat([p.grad.detach().reshape(-1).float() for p in params])
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    mean = torch.stack(gathered).double().mean(0)
    scale = sg.ddp.allreduce_grads(sg.dis_optim)
    got = torch.cat([p.grad.detach().reshape(-1).double() for p in params]) * scale
    err = float((got - mean).norm() / (mean.norm() + 1e-30))
    if not err < 1e-5:
        msg = "allreduce != mean of rank grads (rel %.3e)" % err
    sg.dis_optim.zero_grad()
    # (2) replicas identical after two full steps on rank-specific data
    for i in range(2):
        img = torch.randn(B, 3, R, R, device=device)
        z = torch.randn(B, 512, device=device)
        sg.optimize_discriminator(z, img, depth, alpha)
        sg.optimize_generator(z, img, depth, alpha)
    d = digest([sg.gen, sg.dis, sg.gen_shadow])
    all_d = [torch.empty_like(d) for _ in range(world)]
    dist.all_gather(all_d, d)
    for r in range(1, world):
        if not torch.equal(all_d[0], all_d[r]):
            bad = int((all_d[0] != all_d[r]).nonzero()[0])
            if msg == "ok":
                msg = "replica %d differs from replica 0 after 2 steps (digest entry %d)" % (r, bad)
            break
    flag = torch.tensor([0 if msg == "ok" else 1], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag) and msg == "ok":
        msg = "violation seen on another rank"
    return msg


def run_reference(args, rank, world):
    if rank != 0:
        return
    # bounded sample: one CPU step at 256x256 costs ~12 s per image on a 64-thread host, so a K+W of more than a handful
    # of steps runs at batch 1 and with a single warm-up step (the CPU path has no autotuning or lazy initialisation to
    # absorb); the JSON line reports the warm-up actually run
    batch = args.ref_batch if args.steps + args.warmup <= 4 else 1
    args.warmup = min(args.warmup, 1)
    times, threads = cpu_oracle_step_time(args.res, batch, args.steps, args.warmup,
                                          budget_s=float(os.environ.get("SG_REFERENCE_BUDGET_S", "180")))
    sec = sum(times) / len(times)
    val = batch / sec
    steps_requested, args.steps = args.steps, len(times)   # a slow host stops early (budget): report the steps really timed
    out = {"impl": "reference", "metric": "images/sec G+D step @%dx%d bs=%d/GPU" % (args.res, args.res, args.batch),
           "value": round(val, 4),
           "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": round(sec * 1e3, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": "G+D step (logistic+R1, EMA) %dx%d depth %d" % (args.res, args.res, O_depth(args.res)),
                      "batch_per_step": batch, "device": "host CPU"},
           "cpu_baseline": {"value": round(val, 4), "unit": "images/sec", "cores": threads, "kind": "port",
                            "sample": "%d timed G+D steps (of %d requested; time budget %s s) at batch %d (the reference is "
                                      "Python and absent on the GPU box; this is oracle/stylegan_oracle.py, pinned to it in "
                                      "the build container)"
                                      % (args.steps, steps_requested, os.environ.get("SG_REFERENCE_BUDGET_S", "180"), batch)},
           "e2e": {"value": round(val, 4), "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def O_depth(res):
    import math
    return int(math.log2(res)) - 2


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sgb200")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--ref-batch", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--mode", default="bf16")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.impl == "gpu_baseline":
        t0 = time.time()
        msb = gpu_oracle_step_time(args.res, args.batch, 3, 1, args.mode)
        print(json.dumps({"value": round(args.batch / (msb * 1e-3), 2), "ms_per_step": round(msb, 2),
                          "wall_s_incl_warmup": round(time.time() - t0, 1),
                          "how": "autocast(bfloat16) + channels_last, cudnn.benchmark=True" if args.mode == "bf16"
                                 else "fp32, cudnn.benchmark=True (reference train.py:77)"}))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    if not os.environ.get("SG_BENCH_QUICK"):
        args.warmup = max(args.warmup, 3)
    verbose = bool(os.environ.get("SG_BENCH_VERBOSE"))

    from models import _lib, Losses
    sg = make_gan(args.res, args.precision, device, seed=1234 + rank)
    depth, alpha, B, R = O_depth(args.res), 0.5, args.batch, args.res
    torch.manual_seed(99 + rank)
    n_buf = 4
    imgs_dev = [torch.randn(B, 3, R, R, device=device) for _ in range(n_buf)]
    z_dev = [torch.randn(B, 512, device=device) for _ in range(n_buf)]
    imgs_host = [torch.randn(B, 3, R, R).pin_memory() for _ in range(n_buf)]
    z_host = [torch.randn(B, 512).pin_memory() for _ in range(n_buf)]

    def step_resident(i):
        t0 = time.perf_counter()
        sg.optimize_discriminator(z_dev[i % n_buf], imgs_dev[i % n_buf], depth, alpha)
        t1 = time.perf_counter()
        sg.optimize_generator(z_dev[i % n_buf], imgs_dev[i % n_buf], depth, alpha)
        if verbose and rank == 0:
            print("step %d: D %.3fs G %.3fs  mem %.1f GB" % (i, t1 - t0, time.perf_counter() - t1,
                                                         torch.cuda.max_memory_allocated() / 2 ** 30),
                  file=sys.stderr, flush=True)

    def step_e2e(i):
        img = imgs_host[i % n_buf].to(device, non_blocking=True)
        z = z_host[i % n_buf].to(device, non_blocking=True)
        dl = sg.optimize_discriminator(z, img, depth, alpha)      # returns loss.item(): device->host read each step
        gl = sg.optimize_generator(z, img, depth, alpha)
        return dl, gl

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms)

    import gc
    dp_check = None
    if world > 1:
        dp_check = dp_self_check(sg, device, world, rank, depth, alpha, B, R)
        torch.manual_seed(99 + rank)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                    # spawned before the warm-up so that its start-up cost is not in the timing
    for i in range(args.warmup):
        step_resident(i)
    gc.collect()
    gc.disable()                           # no collector pauses inside the timed regions
    l0 = _lib.launches()
    ms = timed(step_resident, args.steps)
    launches = _lib.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    for i in range(2):                     # untimed: absorbs the sampler's NVML teardown and first-touch of the pinned buffers
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    gc.enable()

    if rank == 0:
        pk, how = peaks()
        value = world * B * args.steps / (ms * 1e-3)
        e2e = world * B * args.steps / (ms_e2e * 1e-3)
        fstep = F_STEP_GFLOP.get(R)
        out = {"metric": "images/sec G+D step @%dx%d bs=%d/GPU" % (R, R, B), "value": round(value, 2),
               "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": args.precision if args.precision != "fp32" else "f32",
               "data": "synthetic",
               "config": {"workload": "StyleGAN G+D step (optimize_discriminator + optimize_generator, logistic + R1, "
                                      "EMA, style mixing, truncation) %dx%d depth %d alpha 0.5" % (R, R, depth),
                          "batch_per_gpu": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                          "shared_real_pass": bool(Losses.LogisticGAN.shared_real_pass),
                          "cuda_graphs": bool(sg.use_cuda_graphs) and any(e.graph is not None for e in sg._graphs.entries.values()),
                          "l2": "per-step activations (GBs) exceed the 126 MB L2; no explicit flush"},
               "e2e": {"value": round(e2e, 2), "unit": "images/sec", "ms_per_step": round(ms_e2e / args.steps, 3),
                       "h2d_bytes_per_step": B * 3 * R * R * 4 + B * 512 * 4, "d2h_bytes_per_step": 8},
               "gpu_launches": int(launches), "clocks": clocks}
        if dp_check is not None:
            out["dp_check"] = dp_check
        if fstep:
            fexec = (F_EXEC_GFLOP if Losses.LogisticGAN.shared_real_pass else F_STEP_GFLOP).get(R)
            out["step_tensor_fraction"] = {
                # EXECUTED FLOPs (4 F_G + 11 F_D with the shared D(real) graph) are what the fraction is computed from
                "executed_gflop_per_image": fexec,
                "achieved_tflops_per_gpu": round(value / world * fexec / 1e3, 1),
                "frac_of_sustained_bf16_peak": round(value / world * fexec / 1e3 / pk["bf16_tflops_sustained"], 4),
                # the reference formulation's count (4 F_G + 14 F_D) for comparison only
                "reference_formulation_gflop_per_image": fstep,
                "reference_equivalent_tflops_per_gpu": round(value / world * fstep / 1e3, 1),
                "peak_source": how}
        ch = {128: 128, 256: 64, 512: 32, 1024: 16}.get(R, 64)
        try:
            out["roofline"] = northstar_kernel_roofline(B, R, ch) if args.precision == "bf16" else None
        except Exception as e:                              # the roofline probe must never lose the headline number
            out["roofline"] = {"error": str(e)[:200]}
        if world == 1 and not args.no_gpu_baseline:
            # honest GPU comparator: eager PyTorch + cuDNN/cuBLAS on this same B200, same step, same batch
            gb = {"unit": "images/sec", "batch": B, "what": "oracle port of the reference step, eager PyTorch on this "
                  "GPU (cuDNN/cuBLAS; none of this library's kernels)", "timing": "1 warm-up + 3 timed steps, CUDA events"}
            del sg
            gc.collect(); torch.cuda.empty_cache()
            for mode in ("bf16", "fp32"):
                # own process + time limit: cudnn.benchmark autotunes every new conv shape (fwd, dgrad, wgrad and their
                # double-backward forms) during the warm-up step, which can take minutes
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "gpu_baseline", "--mode",
                                        mode, "--res", str(R), "--batch", str(B)], capture_output=True, text=True,
                                       timeout=float(os.environ.get("SG_GPU_BASELINE_TIMEOUT", "150")))
                    gb[mode] = json.loads(r.stdout.strip().splitlines()[-1])
                except subprocess.TimeoutExpired:
                    gb[mode] = {"error": "timeout"}
                except Exception as e:
                    gb[mode] = {"error": str(e)[:160]}
            best = max((gb[m].get("value", 0.0) for m in ("bf16", "fp32")), default=0.0)
            gb["speedup_vs_best"] = round(value / best, 2) if best else None
            out["gpu_baseline"] = gb
        if world == 1 and not args.no_cpu_baseline:
            cb = 1                                          # ~10 s of CPU per image at 256x256: 1 warm-up + 3 timed steps
            t, threads = cpu_oracle_step_time(R, cb, 3, 1, budget_s=45.0)
            out["cpu_baseline"] = {"value": round(cb * len(t) / sum(t), 4), "unit": "images/sec", "cores": threads,
                                   "kind": "port", "sample": "1 warm-up + %d timed G+D steps (3 unless the 45 s budget of this leg ran out) at "
                                   "%dx%d, batch %d (the per-image cost of this path does not depend on the batch), oracle "
                                   "port on the host CPU, all %d threads" % (len(t), R, R, cb, threads),
                                   "sec_per_step": [round(x, 2) for x in t]}
        print(json.dumps(out), flush=True)
    if world > 1:
        # Leave without tearing NCCL down: the captured CUDA graphs hold NCCL kernels, and destroy_process_group() with
        # them alive did not return on a 2-GPU box (the JSON line was out, torchrun waited for the ranks until its timeout).
        # A barrier keeps the ranks together until rank 0 has printed; a watchdog bounds even that.
        import threading
        sys.stdout.flush(); sys.stderr.flush()
        threading.Timer(60.0, lambda: os._exit(0)).start()
        try:
            torch.cuda.synchronize()
            dist.barrier()
        except Exception:
            pass
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
