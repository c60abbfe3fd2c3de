"""gpurun_out/parity/parity.jsonl (written by tests/layerwise.py during `pytest -m gpu`) -> profiles/r02_parity.md.

    python tools/parity_table.py [in.jsonl] [out.md]

One row per measurement (the last record of a name wins).  Columns are relative L2 errors of the CUDA path against the
teacher-forced oracle: worst stored tensor, worst / median variable gradient, input gradient, worst loss, worst
post-step weight update."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fmt(x):
    return "—" if x is None else f"{x:.1e}"


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity", "parity.jsonl")
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "r02_parity.md")
    recs = {}
    for line in open(src):
        r = json.loads(line)
        recs[r["name"]] = r
    rows, per_layer = [], []
    for name in sorted(recs):
        r = recs[name]
        if "errors" in r:                                   # free-running record: a dict of named errors
            rows.append((name, r["mode"], None, "", max(r["errors"].values()), None, None, None, None, "free running: " +
                         ", ".join(f"{k} {v:.1e}" for k, v in r["errors"].items())))
            continue
        if "layers" not in r:                               # free/...: {label: {y, dx, grad_max, grad_median}}
            note = "; ".join(f"{k}: y {v['y']:.1e} dx {v['dx']:.1e} grad max {v['grad_max']:.1e} median {v['grad_median']:.1e}"
                             for k, v in r.items() if isinstance(v, dict))
            rows.append((name, r["mode"], None, "", None, None, None, None, None, "free running (reported): " + note))
            continue
        if isinstance(r["layers"], dict):                   # whole train step
            layers = [(k, t, op, e) for k, v in r["layers"].items() for t, op, e in v]
            grads = [e for v in r["grads"].values() for e in v]
            worst = max(layers, key=lambda x: x[3])
            rows.append((name, r["mode"], worst[3], f"{worst[0]}:{worst[2]}#{worst[1]}", max(grads), float(np.median(grads)),
                         None, max(r["metrics"].values()), max(r["update"].values()) if r.get("update") else None,
                         "; ".join(f"{n} {max(v):.1e} ({len(v)} vars)" for n, v in r["grads"].items())))
            if name.startswith("step/C3/256x1") or name.startswith("step/C2/256x1"):
                per_layer.append((name, r))
        else:                                               # single net
            worst = max(r["layers"], key=lambda x: x[2])
            rows.append((name, r["mode"], worst[2], f"{worst[1]}#{worst[0]}", max(r["grads"]), float(np.median(r["grads"])),
                         r["dx"], None, None, f"{len(r['grads'])} vars"))
    out = ["# Parity measurements of round 2 (CUDA path vs the teacher-forced oracle)", "",
           "Written by `tools/parity_table.py` from the records `tests/layerwise.py` appends during `pytest -m gpu` on a B200.",
           "Every number is a relative L2 error. Gates: fp32 check mode 1e-4 (layers, losses, gradients); bf16 mode one bf16 ulp",
           "(3.9e-3) per stored tensor and 2e-2 per variable gradient / loss; `kernel/*` (bf16-representable data against fp64):",
           "6e-3 for kernel and input gradients, 1e-2 for per-channel vectors.", "",
           "| measurement | mode | worst stored tensor | where | worst gradient | median gradient | input gradient | worst loss | worst weight update | detail |",
           "|---|---|---:|---|---:|---:|---:|---:|---:|---|"]
    for n, mode, wl, where, wg, mg, dx, ml, up, note in rows:
        out.append(f"| `{n}` | {mode} | {fmt(wl)} | {where} | {fmt(wg)} | {fmt(mg)} | {fmt(dx)} | {fmt(ml)} | {fmt(up)} | {note} |")
    for name, r in per_layer:
        out += ["", f"## Per-variable gradient errors, `{name}`", ""]
        for net, v in r["grads"].items():
            out.append(f"* **{net}** ({len(v)} variables, Keras `trainable_variables` order): " + " ".join(f"{e:.1e}" for e in v))
        out += ["", "Worst stored tensor per model call: " +
                ", ".join(f"{k} {max(e for _, _, e in v):.1e}" for k, v in r["layers"].items())]
    with open(dst, "w") as fh:
        fh.write("\n".join(out) + "\n")
    print(dst, len(rows), "rows")


if __name__ == "__main__":
    main()
