#!/usr/bin/env python
"""Benchmark of the CycleGAN train step (BASELINE.json metric: train images/sec at 256x256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C2|C2s|C1|C5] [--mode bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...       # the CPU restatement of the reference (torch-CPU oracle), host cores

One "step" = one full CycleGan.train_step (6 generator forwards, 4 discriminator forwards, the combined backward,
4 Adam updates) on one synthetic batch.  `value` = (A,B) pairs per second over all ranks with the inputs already in
HBM; `e2e` = the same through the public API with pinned HOST inputs (H2D inside the timed region) and a D2H read
of the 6 metrics every step.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

RESNET64 = dict(type="resnet_generator", filters=64)
SIMPLE_D4 = dict(type="simple_discriminator", filters=[64, 128, 256, 512], kernels=[4, 4, 4, 4],
                 normalization="instancenorm")
SIMPLE_D3 = dict(type="simple_discriminator", filters=[64, 128, 256], kernels=[4, 4, 4], normalization="instancenorm")
UNET_G = dict(type="unet_generator", filters=[16, 32, 64, 128], kernels=[4, 4, 4, 4], output_channels=3,
              expansion="upsample", normalization="instancenorm", dropout=False, final_activation="tanh")
UNET_D = dict(type="unet_generator", filters=[16, 32, 64], kernels=[7, 5, 3], output_channels=1,
              expansion="upsample", normalization="instancenorm", dropout=False, final_activation="sigmoid")
STRIDED7 = dict(type="strided_unet", filters=[64, 128, 256, 512, 512, 512, 512], kernels=[4] * 7, output_channels=3,
                normalization="instancenorm", final_activation="tanh")
WORKLOADS = {
    "C1": dict(gen=UNET_G, disc=SIMPLE_D3, size=128, batch=1, name="C1: unet_generator(cycle.yaml) + simple_discriminator[64,128,256], 128x128"),
    "C2": dict(gen=UNET_G, disc=UNET_D, size=256, batch=4, name="C2: cycle.yaml verbatim (U-Net G + U-Net PatchGAN D), 256x256"),
    "C2s": dict(gen=STRIDED7, disc=UNET_D, size=256, batch=4, name="C2s: strided_unet-7 G + U-Net PatchGAN D, 256x256"),
    "C3": dict(gen=RESNET64, disc=SIMPLE_D4, size=256, batch=16, name="C3: resnet_generator{filters:64} (9 blocks) + simple_discriminator[64,128,256,512] k4, 256x256"),
    # BASELINE.json configs[4]: generator inference (predict.py path), replicas only -- no discriminator, no backward
    "C5": dict(gen=UNET_G, disc=None, size=512, batch=32, name="C5: unet_generator(cycle.yaml) forward only (predict.py path), 512x512"),
}
LOSS_WEIGHTS = dict(cycle=2.0, identity=0.5, generator=1.0, discriminator=0.5)
ADAM = dict(name="adam", learning_rate=2e-4, beta_1=0.5)
METRIC = "cyclegan_train_pairs_per_sec_256x256"
UNIT = "images/s"


def synthetic_batch(batch, size, rank=0):
    a = np.random.RandomState(1234 + 2 * rank).uniform(-1.0, 1.0, size=(batch, size, size, 3)).astype(np.float32)
    b = np.random.RandomState(1235 + 2 * rank).uniform(-1.0, 1.0, size=(batch, size, size, 3)).astype(np.float32)
    return a, b


def step_flops(gen_graph, disc_graph, size):
    """Algorithmic FLOPs per (a,b) pair: 18 F_G + 14 F_D - 4 f_G0 - 4 f_D0."""
    return 18 * gen_graph.flops(size, size) + 14 * disc_graph.flops(size, size) \
        - 4 * gen_graph.first_layer_flops(size, size) - 4 * disc_graph.first_layer_flops(size, size)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), hbm=float(d["hbm_gbs"]),
                    source="measured (MEASURED_PEAKS.json, bf16_tflops_sustained: kernel timed inside a long step)")
    return dict(tflops=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                     "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80}
            while not self._halt.is_set():
                self.samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
                time.sleep(0.1)
        except Exception as e:       # NVML missing: record that instead of failing the bench
            self.reasons.add(f"nvml_unavailable:{type(e).__name__}")

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return dict(sm_mhz=(s[len(s) // 2] if s else None), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


def bunch(**kw):
    from cyclegan_cat_b200.model_processing.load_model import Bunch
    return Bunch(**kw)


# ---------------------------------------------------------------------------------------------
# CPU arm: the restated reference (oracle) on the host cores
# ---------------------------------------------------------------------------------------------
def time_oracle(wl, steps, warmup, sample_batch=1, budget_s=150.0):
    import torch
    from oracle.train import OracleCycleGan
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = OracleCycleGan(wl["gen"], wl["disc"], loss="mse", loss_weights=LOSS_WEIGHTS, g_opt=ADAM, d_opt=ADAM)
    a, b = synthetic_batch(sample_batch, wl["size"])
    t0 = time.perf_counter()
    o.train_step(a, b)                      # first step also pays one-time oneDNN primitive creation
    first = time.perf_counter() - t0
    warm_done = 1
    while warm_done < warmup and (time.perf_counter() - t0) < budget_s * 0.3:
        o.train_step(a, b)
        warm_done += 1
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t1 = time.perf_counter()
        o.train_step(a, b)
        times.append(time.perf_counter() - t1)
        if time.perf_counter() - t_start > budget_s:
            break
    med = float(np.median(times))
    return dict(value=sample_batch / med, ms_per_step=med * 1e3, cores=cores, steps_done=len(times),
                warmup_done=warm_done, first_step_s=first,
                sample=f"{wl['name']}, batch {sample_batch}, {len(times)} timed steps (median), fp32 torch-CPU "
                       f"restatement of the reference (oneDNN), not TensorFlow")


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_oracle(wl, args.steps, args.warmup)
    line = dict(metric=METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=r["steps_done"],
                warmup=r["warmup_done"], ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=dict(workload=wl["name"], image_size=wl["size"], batch_per_step=1,
                            note="reference = torch-CPU restatement of cyclegan/model.py:136-154 (TensorFlow is not "
                                 "installable here); one step = one train_step on a batch-1 sample of the workload"),
                cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"]),
                e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# C5: the predict.py path (predict.py:20-36) -- uint8 image -> normalize -> generator -> uint8, replicas only
# ---------------------------------------------------------------------------------------------
INFER_METRIC = "cyclegan_predict_images_per_sec_512x512"


def time_oracle_forward(wl, steps, budget_s=60.0):
    import torch
    from oracle import models as om, tf_ops as T
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    o = om.create_model(wl["gen"])
    o.load(om.init_variables(o.var_specs, 42))
    u8 = np.random.RandomState(1234).randint(0, 256, size=(1, wl["size"], wl["size"], 3)).astype(np.uint8)

    def once():
        with torch.no_grad():
            return T.postprocess_prediction(o(T.normalize(u8)).numpy())
    once()
    times, t0 = [], time.perf_counter()
    for _ in range(steps):
        t1 = time.perf_counter()
        once()
        times.append(time.perf_counter() - t1)
        if time.perf_counter() - t0 > budget_s:
            break
    med = float(np.median(times))
    return dict(value=1.0 / med, ms_per_step=med * 1e3, cores=cores, steps_done=len(times),
                sample=f"{wl['name']}, batch 1, {len(times)} timed calls (median): normalize -> generator -> uint8, fp32 "
                       f"torch-CPU restatement of the reference (oneDNN), not TensorFlow")


def run_infer_reference(args, wl):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    r = time_oracle_forward(wl, max(args.steps, 3))
    print(json.dumps(dict(metric=INFER_METRIC, value=r["value"], unit=UNIT, n_gpus=args.gpus, steps=r["steps_done"],
                          warmup=1, ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak", vs_baseline=None,
                          dtype="f32", data="synthetic", impl="reference",
                          config=dict(workload=wl["name"], image_size=wl["size"], batch_per_step=1),
                          cpu_baseline=dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"]),
                          e2e=dict(value=r["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))), flush=True)


def run_infer(args, wl):
    import ctypes
    import torch
    import torch.distributed as dist
    from cyclegan_cat_b200 import _lib
    from cyclegan_cat_b200.cyclegan.model import create_model
    from cyclegan_cat_b200.transform import data_load as DL

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, S = args.batch or wl["batch"], wl["size"]
    g = create_model(wl["gen"], mode=args.mode)
    g.initialize(42)
    lib = _lib.load()
    u8 = np.random.RandomState(1234 + rank).randint(0, 256, size=(B, S, S, 3)).astype(np.uint8)
    u8_pin = torch.from_numpy(u8).pin_memory()
    x_dev = DL.normalize_device(u8).torch
    out_pin = torch.empty((B, S, S, 3), dtype=torch.uint8).pin_memory()
    out_dev = torch.empty((B, S, S, 3), dtype=torch.uint8, device="cuda")
    from cyclegan_cat_b200.runtime import _ptr, _stream_ptr

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def e2e_once():      # predict.py:20-36 for a batch: host uint8 in, host uint8 out
        xd = u8_pin.cuda(non_blocking=True)
        y = g(DL.normalize_device(xd))
        _lib.check(lib.cg_postprocess_u8(_ptr(y.torch), _ptr(out_dev), out_dev.numel(), _stream_ptr(torch)), "post")
        out_pin.copy_(out_dev, non_blocking=True)

    W = max(args.warmup, 3)
    for _ in range(W):
        g(x_dev)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.cg_launch_count(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_range = os.environ.get("CG_PROFILE_STEP") == "1"      # ncu --profile-from-start off: capture exactly the timed calls
    if prof_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        g(x_dev)
    e1.record()
    if prof_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop()
    launches = ctypes.c_int64()
    lib.cg_launch_count(ctypes.byref(launches), 0)
    # tensor-core share: one extra, instrumented call (CUDA-event pairs around every tensor-core launch)
    keep = os.environ.get("CG_KEEP_PROF")             # per-launch CSV of that call (tools/prof_layers.py reads it)
    if keep:
        os.environ["CG_PROF_DUMP"] = keep
    lib.cg_prof_enable(1)
    g(x_dev)
    pms, pl, pfl = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
    lib.cg_prof_read(ctypes.byref(pms), ctypes.byref(pl), ctypes.byref(pfl))
    lib.cg_prof_enable(0)
    os.environ.pop("CG_PROF_DUMP", None)
    for _ in range(2):
        e2e_once()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        e2e_once()
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    ms_step = ms_total / args.steps
    fl_img = g.graph.flops(S, S)
    tfl = fl_img * B / (ms_step * 1e-3) / 1e12
    tc_tfl = (pfl.value / (pms.value * 1e-3) / 1e12) if pms.value > 0 else None
    line = dict(metric=INFER_METRIC, value=world * B * args.steps / (ms_total * 1e-3), unit=UNIT, n_gpus=world,
                steps=args.steps, warmup=W, ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype=args.mode, data="synthetic",
                config=dict(workload=wl["name"], image_size=S, batch_per_gpu=B, parallelism=f"replicas x{world} (no collective)",
                            flops_per_image=fl_img, l2="no flush needed: one call streams > 10 GB of activations (>> 126 MB L2)",
                            weights="random init N(0,0.02), seed 42"),
                roofline=dict(bound="tensor", kernel="conv_tc_kernel<BK16> (16-channel-group tcgen05 conv; bound by the TMA row "
                              "rate of its 32-byte rows, DESIGN.md 3.1), all tensor-core launches of one call",
                              achieved=tc_tfl, peak=pk["tflops"], unit="TFLOP/s", frac=(tc_tfl / pk["tflops"]) if tc_tfl else None,
                              traffic=None, launches=int(pl.value), share_of_step=(pms.value / ms_step) if ms_step > 0 else None,
                              peak_source=pk["source"], whole_call_tflops=tfl, whole_call_frac=tfl / pk["tflops"]),
                clocks=clocks, gpu_launches=int(launches.value),
                e2e=dict(value=world * B * args.steps / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=int(u8.nbytes),
                         d2h_bytes_per_step=int(u8.nbytes), ms_per_step=ms_e2e / args.steps,
                         path="pinned uint8 -> H2D -> cg_normalize_u8 -> cg_net_forward -> cg_postprocess_u8 -> D2H (predict.py:20-36)"))
    if not args.no_cpu_baseline:
        r = time_oracle_forward(wl, 5, budget_s=40.0)
        line["cpu_baseline"] = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"],
                                    ms_per_step=r["ms_per_step"])
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------
# short loops attached to the default line: BASELINE.json configs[1] (C2), configs[4] (C5), and C3 at the per-GPU batch of
# configs[3] (global batch 64 split over the ranks: the strong-scaling point next to the weak headline)
# ---------------------------------------------------------------------------------------------
# HBM-side roofline of the narrow-N U-Net workloads: per conv-output element the forward must at least write
# it (conv), read and re-write it (instance norm + activation) and read it again (next conv) = 8 bytes in bf16; the backward
# reads x and dy for the norm's reduction and again for its apply, writes the norm's dx, reads it (weight gradient) and
# writes / reads the conv's dx = 16 bytes.  Train step: 6 generator + 4 discriminator calls forward, all of them backward
# (the discriminators twice on the fake half) -> 24 bytes x elements x calls; inference: 8 bytes x elements.
def hbm_bytes_train(gen_graph, disc_graph, size, batch):
    eg, ed = gen_graph.conv_out_elems(size, size), disc_graph.conv_out_elems(size, size)
    return batch * (6 * eg * 24 + 4 * ed * 24 + 2 * ed * 8)


def quick_train(wl, mode, batch, steps, warmup, world, barrier, max_over_ranks, rank):
    """Device-resident and end-to-end train-step timing of one workload (no per-launch events)."""
    import torch
    from cyclegan_cat_b200.cyclegan.model import CycleGan
    S = wl["size"]
    mc = bunch(name="bench", new=True, location="/tmp/cg_b200_bench", generator=dict(wl["gen"]),
               discriminator=dict(wl["disc"]), loss="mse", loss_weights=dict(LOSS_WEIGHTS))
    tc = bunch(epochs=1, batch_size=batch, image_size=S, g_opt=dict(ADAM), d_opt=dict(ADAM),
               summary=dict(samples=1, images=5, model=20))
    gan = CycleGan(mc, tc, mode=mode)
    for i, n in enumerate((gan.g_AB, gan.g_BA, gan.d_A, gan.d_B)):
        n.initialize(42 + i)
    gan.prepare(batch, S, S)
    if world > 1:
        gan.enable_data_parallel()
    a_np, b_np = synthetic_batch(batch, S, rank)
    a_dev, b_dev = torch.from_numpy(a_np).cuda(), torch.from_numpy(b_np).cuda()
    a_pin, b_pin = torch.from_numpy(a_np).pin_memory(), torch.from_numpy(b_np).pin_memory()
    for _ in range(max(warmup, 3)):
        gan.train_step(a_dev, b_dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        gan.train_step(a_dev, b_dev)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    for _ in range(2):
        gan.train_step(a_pin, b_pin)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(steps):
        mm = gan.train_step(a_pin, b_pin)
        _ = [float(v) for v in mm.values()]
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    fl = step_flops(gan.g_AB.graph, gan.d_A.graph, S) * batch
    hb = hbm_bytes_train(gan.g_AB.graph, gan.d_A.graph, S, batch)
    pk = peaks()
    ms_step = ms / steps
    out = dict(workload=wl["name"], batch_per_gpu=batch, global_batch=batch * world, steps=steps,
               value=world * batch * steps / (ms * 1e-3), unit=UNIT, ms_per_step=ms_step,
               e2e=dict(value=world * batch * steps / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=int(2 * a_np.nbytes),
                        d2h_bytes_per_step=24, ms_per_step=ms_e2e / steps),
               roofline=dict(tensor=dict(achieved=fl / (ms_step * 1e-3) / 1e12, peak=pk["tflops"], unit="TFLOP/s",
                                         frac=fl / (ms_step * 1e-3) / 1e12 / pk["tflops"], flops_per_step=fl),
                             hbm=dict(achieved=hb / (ms_step * 1e-3) / 1e9, peak=pk["hbm"], unit="GB/s",
                                      frac=hb / (ms_step * 1e-3) / 1e9 / pk["hbm"], algorithmic_bytes_per_step=hb,
                                      model="24 B per conv-output element and model call (bench.py hbm_bytes_train)"),
                             note="whole-step figures: algorithmic FLOPs / bytes of the step over its measured time"))
    del gan
    torch.cuda.empty_cache()
    return out


def quick_infer(wl, mode, batch, steps, warmup, world, barrier, max_over_ranks, rank):
    """Generator inference (predict.py path): device-resident calls and the uint8-to-uint8 end-to-end loop."""
    import torch
    from cyclegan_cat_b200 import _lib
    from cyclegan_cat_b200.cyclegan.model import create_model
    from cyclegan_cat_b200.runtime import _ptr, _stream_ptr
    from cyclegan_cat_b200.transform import data_load as DL
    S = wl["size"]
    lib = _lib.load()
    g = create_model(wl["gen"], mode=mode)
    g.initialize(42)
    u8 = np.random.RandomState(1234 + rank).randint(0, 256, size=(batch, S, S, 3)).astype(np.uint8)
    u8_pin = torch.from_numpy(u8).pin_memory()
    x_dev = DL.normalize_device(u8).torch
    out_pin = torch.empty((batch, S, S, 3), dtype=torch.uint8).pin_memory()
    out_dev = torch.empty((batch, S, S, 3), dtype=torch.uint8, device="cuda")

    def e2e_once():
        xd = u8_pin.cuda(non_blocking=True)
        y = g(DL.normalize_device(xd))
        _lib.check(lib.cg_postprocess_u8(_ptr(y.torch), _ptr(out_dev), out_dev.numel(), _stream_ptr(torch)), "post")
        out_pin.copy_(out_dev, non_blocking=True)
    for _ in range(max(warmup, 3)):
        g(x_dev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g(x_dev)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    for _ in range(2):
        e2e_once()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(steps):
        e2e_once()
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    pk = peaks()
    ms_step = ms / steps
    fl = g.graph.flops(S, S) * batch
    hb = g.graph.conv_out_elems(S, S) * 8 * batch
    out = dict(workload=wl["name"], metric=INFER_METRIC, batch_per_gpu=batch, steps=steps,
               value=world * batch * steps / (ms * 1e-3), unit=UNIT, ms_per_step=ms_step,
               e2e=dict(value=world * batch * steps / (ms_e2e * 1e-3), unit=UNIT, h2d_bytes_per_step=int(u8.nbytes),
                        d2h_bytes_per_step=int(u8.nbytes), ms_per_step=ms_e2e / steps),
               roofline=dict(tensor=dict(achieved=fl / (ms_step * 1e-3) / 1e12, peak=pk["tflops"], unit="TFLOP/s",
                                         frac=fl / (ms_step * 1e-3) / 1e12 / pk["tflops"], flops_per_step=fl),
                             hbm=dict(achieved=hb / (ms_step * 1e-3) / 1e9, peak=pk["hbm"], unit="GB/s",
                                      frac=hb / (ms_step * 1e-3) / 1e9 / pk["hbm"], algorithmic_bytes_per_step=hb,
                                      model="8 B per conv-output element (conv write, norm read + write, next conv read)")))
    del g
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args, wl):
    import ctypes
    import torch
    import torch.distributed as dist
    from cyclegan_cat_b200 import _lib
    from cyclegan_cat_b200.cyclegan.model import CycleGan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B, S = args.batch or wl["batch"], wl["size"]

    mc = bunch(name="bench", new=True, location="/tmp/cg_b200_bench", generator=dict(wl["gen"]),
               discriminator=dict(wl["disc"]), loss="mse", loss_weights=dict(LOSS_WEIGHTS))
    tc = bunch(epochs=1, batch_size=B, image_size=S, g_opt=dict(ADAM), d_opt=dict(ADAM),
               summary=dict(samples=1, images=5, model=20))
    gan = CycleGan(mc, tc, mode=args.mode)
    for i, n in enumerate((gan.g_AB, gan.g_BA, gan.d_A, gan.d_B)):
        n.initialize(42 + i)                                  # SURVEY 8d seeds
    gan.prepare(B, S, S)
    if world > 1:
        gan.enable_data_parallel()
    lib = _lib.load()

    a_np, b_np = synthetic_batch(B, S, rank)
    a_dev, b_dev = torch.from_numpy(a_np).cuda(), torch.from_numpy(b_np).cuda()
    a_pin, b_pin = torch.from_numpy(a_np).pin_memory(), torch.from_numpy(b_np).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput -----------------------------------------------------------
    per_step = ctypes.c_int64()
    for i in range(max(args.warmup, 3)):
        if i == 1:
            lib.cg_prof_enable(1)                  # count the profiled launches of one step ...
        gan.train_step(a_dev, b_dev)
        if i == 1:
            lib.cg_prof_read(None, ctypes.byref(per_step), None)
            lib.cg_prof_enable(0)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lib.cg_launch_count(None, 1)
    prof_csv = os.path.join("/tmp", f"cg_tc_launches_{os.getpid()}.csv")
    os.environ["CG_PROF_DUMP"] = prof_csv          # per-launch (geometry, flops, CUDA-event ms) of the tensor-core kernels
    # The per-launch CUDA events of the tensor-core kernels (two per launch) cost ~1.2 ms of a 48 ms step, so they are
    # recorded in ONE of the K timed steps (the middle one); the other steps run uninstrumented.
    prof_on = os.environ.get("CG_BENCH_NO_PROF") != "1"
    prof_step = args.steps // 2
    if prof_on:
        lib.cg_prof_enable(int(per_step.value) + 16)   # ... create the CUDA events before the timed region ...
        lib.cg_prof_enable(-1)                         # ... and pause until the profiled step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    prof_range = os.environ.get("CG_PROFILE_STEP") == "1"      # ncu --profile-from-start off: capture exactly the timed steps
    if prof_range:
        torch.cuda.profiler.start()
    e0.record()
    for k in range(args.steps):
        if prof_on and k == prof_step:
            lib.cg_prof_enable(1)
        m = gan.train_step(a_dev, b_dev)
        if prof_on and k == prof_step:
            lib.cg_prof_enable(-1)
    e1.record()
    if prof_range:
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    ms_mine = e0.elapsed_time(e1)
    ms_total = max_over_ranks(ms_mine)
    rank_ms = [ms_mine / args.steps]
    if world > 1:       # per-rank step time of the same timed region: the spread is the rank skew max_over_ranks pays for
        t_all = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(t_all, torch.tensor([ms_mine / args.steps], dtype=torch.float64, device="cuda"))
        rank_ms = [float(t.item()) for t in t_all]
    clocks = sampler.stop()
    launches = ctypes.c_int64()
    lib.cg_launch_count(ctypes.byref(launches), 0)
    pms, pl, pfl = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
    lib.cg_prof_read(ctypes.byref(pms), ctypes.byref(pl), ctypes.byref(pfl))
    lib.cg_prof_enable(0)
    last_metrics = {k: float(v) for k, v in m.items()}

    # ---- end to end through the public API: pinned host inputs, metrics read back every step ---------
    for _ in range(0 if args.no_e2e else 2):
        gan.train_step(a_pin, b_pin)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(1 if args.no_e2e else args.steps):
        mm = gan.train_step(a_pin, b_pin)
        _ = [float(v) for v in mm.values()]                  # D2H of the 6 metrics (what model.py:301 does)
    e3.record()
    barrier()
    ms_e2e = max_over_ranks(e2.elapsed_time(e3))

    # ---- attached workloads (default C3 invocation only; every rank takes part, rank 0 reports) ------------------
    extra = {}
    g_graph, d_graph = gan.g_AB.graph, gan.d_A.graph
    if args.workload == "C3" and not args.batch and not args.no_extra and args.mode == "bf16":
        del gan
        torch.cuda.empty_cache()
        q = dict(mode=args.mode, warmup=3, world=world, barrier=barrier, max_over_ranks=max_over_ranks, rank=rank)
        if 64 % world == 0:      # BASELINE.json configs[3]: global batch 64 over the ranks (strong scaling; at N = 1 its baseline)
            extra["c4_strong"] = quick_train(WORKLOADS["C3"], batch=64 // world, steps=4 if world == 1 else 8, **q)
            extra["c4_strong"]["scaling"] = "strong"
        extra["C2"] = quick_train(WORKLOADS["C2"], batch=WORKLOADS["C2"]["batch"], steps=10, **q)
        extra["C5"] = quick_infer(WORKLOADS["C5"], batch=WORKLOADS["C5"]["batch"], steps=5, **q)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    # dominant kernel = conv_tc_kernel at the geometry with the largest share of the timed region
    dom = dict(ms=0.0, flops=0.0, n=0, geom=None)
    try:
        import csv
        agg = {}
        for r in csv.DictReader(open(prof_csv)):
            if int(r["kind"]) not in (1, 2, 5):          # conv_tc_kernel: 64-channel form, pair kernel, window form
                continue
            k = (int(r["taps"]), int(r["cchunks"]), int(r["bn"]))
            a = agg.setdefault(k, [0.0, 0.0, 0])
            a[0] += float(r["ms"]); a[1] += float(r["flops"]); a[2] += 1
        if agg:
            k, a = max(agg.items(), key=lambda kv: kv[1][0])
            dom = dict(ms=a[0], flops=a[1], n=a[2], geom=f"taps={k[0]} k_chunks={k[1]} n_tile={k[2]}")
        keep = os.environ.get("CG_KEEP_PROF")             # copy the per-launch CSV somewhere (tools/prof_layers.py reads it)
        if keep:
            import shutil
            shutil.copy(prof_csv, keep)
        os.remove(prof_csv)
    except Exception as e:      # the aggregate below is still reported
        dom["geom"] = f"unavailable ({type(e).__name__})"
    # dram bytes per launch from the committed `ncu --set full` capture: valid only for the geometry it was taken at (the C3
    # trunk conv, 32-image launch); any other dominant kernel reports null rather than a number measured elsewhere
    traffic, traffic_source = None, None
    tp = os.path.join(ROOT, "profiles", "r01_tc_traffic.json")
    if os.path.exists(tp) and args.workload == "C3" and B == 16 and dom["geom"] == "taps=9 k_chunks=4 n_tile=256":
        traffic = json.load(open(tp)).get("conv_tc_kernel_trunk_fwd_dram_bytes_per_launch")
        traffic_source = "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one 32-image trunk launch (profiles/r01_ncu_trunk_conv.md)"
    fl_pair = step_flops(g_graph, d_graph, S)
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    step_tflops = fl_pair * B / (ms_step * 1e-3) / 1e12
    roof = dict(bound="tensor", kernel=f"conv_tc_kernel (tcgen05 implicit-GEMM conv, forward + data gradient) at {dom['geom']}",
                achieved=(dom["flops"] / (dom["ms"] * 1e-3) / 1e12) if dom["ms"] > 0 else None, peak=pk["tflops"],
                unit="TFLOP/s", frac=None, traffic=traffic, traffic_source=traffic_source, launches=int(dom["n"]),
                avg_launch_ms=(dom["ms"] / dom["n"]) if dom["n"] else None,
                share_of_step=(dom["ms"] / ms_step) if ms_total > 0 else None, peak_source=pk["source"],
                algorithmic_flops_per_launch=(dom["flops"] / dom["n"]) if dom["n"] else None,
                all_tensor_core_kernels=dict(achieved=(pfl.value / (pms.value * 1e-3) / 1e12) if pms.value > 0 else None,
                                             launches=int(pl.value), share_of_step=(pms.value / ms_step) if ms_total > 0 else None),
                events=f"CUDA-event pairs around every tensor-core launch of timed step {prof_step + 1} of {args.steps}",
                whole_step_tflops=step_tflops, whole_step_frac=step_tflops / pk["tflops"])
    if roof["achieved"] is not None:
        roof["frac"] = roof["achieved"] / pk["tflops"]
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.mode,
                data="synthetic",
                config=dict(workload=wl["name"], image_size=S, batch_per_gpu=B, global_batch=B * world,
                            parallelism=f"dp{world}", pairs_per_step=B * world, rank_ms_per_step=rank_ms,
                            individual_images_per_sec=2 * value, flops_per_pair=fl_pair,
                            l2="no flush needed: one step streams tens of GB of activations (>> 126 MB L2)",
                            weights="random init N(0,0.02), seeds 42-45", last_metrics=last_metrics),
                roofline=roof, clocks=clocks, gpu_launches=int(launches.value),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=int(2 * a_np.nbytes),
                         d2h_bytes_per_step=24, ms_per_step=ms_e2e / args.steps))
    if not args.no_cpu_baseline:
        c1 = WORKLOADS[args.workload]
        r = time_oracle(c1, steps=8, warmup=1, budget_s=20.0)       # a bounded sample: ~10-20 s of CPU work (8 batch-1 steps of C3)
        line["cpu_baseline"] = dict(value=r["value"], unit=UNIT, cores=r["cores"], kind="port", sample=r["sample"],
                                    ms_per_step=r["ms_per_step"])
    if extra:
        line["workloads"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-input loop (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the attached C2 / C5 / c4_strong loops of the default C3 run")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if wl["disc"] is None:
        (run_infer_reference if args.impl == "reference" else run_infer)(args, wl)
    elif args.impl == "reference":
        run_reference(args, wl)
    else:
        run_gpu(args, wl)


if __name__ == "__main__":
    main()
