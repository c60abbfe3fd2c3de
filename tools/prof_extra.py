"""Achieved HBM bandwidth of the streaming kernels of the optional paths / input pipeline (kernels_extra.cu), CUDA-event
timed on the launching stream after warm-up, working sets larger than the 126 MB L2.  Writes a markdown table.

    python tools/prof_extra.py [out.md]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cyclegan_cat_b200 import _lib, ir                                       # noqa: E402
from cyclegan_cat_b200.cyclegan.model import CycleGan                        # noqa: E402
from cyclegan_cat_b200.model_processing.load_model import Bunch              # noqa: E402
from cyclegan_cat_b200.runtime import Model, _ptr, _stream_ptr               # noqa: E402
from cyclegan_cat_b200.transform import data_load as DL                      # noqa: E402


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "extra_kernels.md")
    lib = _lib.load()
    _lib.init_device(0)
    st = lambda: _stream_ptr(torch)
    peak = 6521.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    rows = []

    def add(name, ms, nbytes, note=""):
        gbs = nbytes / (ms * 1e-3) / 1e9
        rows.append((name, ms * 1e3, nbytes / 1e6, gbs, gbs / peak, note))

    # input pipeline at the C5 batch x2 (64 x 512 x 512 x 3)
    N, S = 64, 512
    u8 = torch.randint(0, 256, (N, S, S, 3), dtype=torch.uint8, device="cuda")
    f32 = torch.empty((N, S, S, 3), dtype=torch.float32, device="cuda")
    n = u8.numel()
    add("normalize_u8_kernel", timed(lambda: lib.cg_normalize_u8(_ptr(u8), _ptr(f32), n, st())), 5 * n, "1 B read + 4 B written / element")
    f32.uniform_(-1, 1)
    add("postprocess_u8_kernel", timed(lambda: lib.cg_postprocess_u8(_ptr(f32), _ptr(u8), n, st())), 5 * n, "4 B read + 1 B written / element")
    src = torch.empty((32, 572, 572, 3), dtype=torch.float32, device="cuda").uniform_(-1, 1)
    dst = torch.empty((32, 512, 512, 3), dtype=torch.float32, device="cuda")
    add("resize_kernel (572^2 -> 512^2)", timed(lambda: lib.cg_resize_bilinear(_ptr(src), 32, 572, 572, 3, _ptr(dst), 512, 512, st())),
        4 * (src.numel() + dst.numel()), "source read once + destination written (algorithmic)")
    oy = torch.randint(0, 51, (32,), dtype=torch.int32, device="cuda")
    fl = torch.randint(0, 2, (32,), dtype=torch.int32, device="cuda")
    add("resize_kernel (random_jitter 512^2 -> 562^2 -> crop 512^2)",
        timed(lambda: lib.cg_resize_crop_flip(_ptr(dst), 32, 512, 512, 3, 562, 562, _ptr(f32), 512, 512, _ptr(oy), _ptr(oy), _ptr(fl), st())),
        4 * 2 * dst.numel(), "<= source read once + destination written")
    # dropout on a bf16-mode activation-sized tensor (the graph converts fp32 <-> bf16 around it, so time the fp32 net)
    g = ir.Graph()
    g.dropout(g.input, 0.5)
    m = Model(g, name="dropout_only", mode="fp32")
    x = torch.empty((32, 512, 512, 3), dtype=torch.float32, device="cuda").uniform_(-1, 1)
    lib.cg_launch_count(None, 1)
    t_id = timed(lambda: m(x))                       # convert in, dropout (copy), convert out
    t_tr = timed(lambda: m(x, training=True))
    add("dropout_kernel<float> fwd inside a 3-kernel call (inference copy)", t_id / 3, 8 * x.numel(), "whole call / 3 launches; 4 B read + 4 B written")
    add("dropout_kernel<float> fwd inside a 3-kernel call (training, hash per element)", t_tr - 2 * t_id / 3, 8 * x.numel(), "call minus the two conversions")
    # optimizers over the C3 parameter set (28.27 M floats in 4 launches)
    gen = dict(type="resnet_generator", filters=64)
    disc = dict(type="simple_discriminator", filters=[64, 128, 256, 512], kernels=[4, 4, 4, 4], normalization="instancenorm")
    for name, bpp in (("adam", 28), ("sgd", 12), ("rmsprop", 20), ("adabelief", 28)):
        opt = dict(name=name, learning_rate=2e-4, beta_1=0.5)
        mc = Bunch(name="p", new=True, location="/tmp/cg_prof_extra", generator=gen, discriminator=disc, loss="mse",
                   loss_weights=dict(cycle=2.0, identity=0.5, generator=1.0, discriminator=0.5))
        tc = Bunch(epochs=1, batch_size=1, image_size=64, g_opt=opt, d_opt=opt, summary=dict(samples=1, images=5, model=20))
        gan = CycleGan(mc, tc, mode="bf16")
        gan.prepare(1, 64, 64)
        for gr in gan._grads:
            gr.normal_(0, 1e-3)
        npar = sum(net.n_params for net in gan._nets())
        add(f"{'adam_kernel' if name == 'adam' else 'opt_kernel<' + name + '>'} (4 launches, {npar / 1e6:.2f} M params)",
            timed(gan.apply_gradients), bpp * npar, f"{bpp} B / parameter")
        del gan
        torch.cuda.empty_cache()
    with open(out, "w") as fh:
        fh.write(f"HBM peak used: {peak:.0f} GB/s (MEASURED_PEAKS.json hbm_gbs or the 6521 fallback).  CUDA events, 20 iterations after 3 warm-ups.\n\n")
        fh.write("| kernel | us / launch(es) | algorithmic MB | achieved GB/s | of peak | note |\n|---|---:|---:|---:|---:|---|\n")
        for r in rows:
            fh.write(f"| `{r[0]}` | {r[1]:.1f} | {r[2]:.1f} | {r[3]:.0f} | {r[4]:.2f} | {r[5]} |\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
