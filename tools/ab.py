#!/usr/bin/env python
"""A/B of library builds on ONE GPU box: runs bench.py once per (spec, repeat), interleaved so
that drift of the box hits every spec alike, and prints scans/s, roofline fraction and kernel ms.

    python tools/ab.py [--repeats 2] [--steps 30] [--tag r2a] [--args "--shape hdl32 --scans 4096"] SPEC...

SPEC = label:variant[:ENV=VALUE,ENV=VALUE...]   variant "" = product library, otherwise
libnsc_b200_<variant>.so (csrc/Makefile VARIANT=...). Labels starting with "x" are measurement-only
builds (no valid descriptors): bench.py runs them with --no-checks. Example:
    tools/ab.py ws: old:tune:NSC_WS=0 d4:d4
"""
import argparse
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("specs", nargs="+")
    ap.add_argument("--repeats", type=int, default=2)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--tag", default="ab")
    ap.add_argument("--args", default="")
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    rows = {}
    for rep in range(a.repeats):
        for spec in a.specs:
            parts = spec.split(":")
            label, variant = parts[0], parts[1] if len(parts) > 1 else ""
            env = dict(os.environ)
            if variant:
                env["NSC_LIB"] = os.path.join(ROOT, "neural_spectral_codec_b200", f"libnsc_b200_{variant}.so")
            if len(parts) > 2 and parts[2]:
                for kv in parts[2].split(","):
                    k, v = kv.split("=", 1)
                    env[k] = v
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(a.steps), "--warmup", "3",
                   "--no-cpu", "--no-extras"] + (["--no-checks"] if label.startswith("x") else []) + a.args.split()
            r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT)
            if r.returncode != 0:
                print(f"{label}: FAILED\n{r.stderr[-1500:]}", flush=True)
                continue
            d = json.loads(r.stdout.strip().splitlines()[-1])
            rows.setdefault(label, []).append(d)
            print(f"{label:16s} rep {rep}  {d['value'] / 1e6:7.3f} M scans/s  frac {d['roofline']['frac']:.4f}  "
                  f"kernel {d['roofline']['kernel_ms']:.4f} ms  e2e {d['e2e']['value']:.0f}", flush=True)
    summary = {}
    for label, ds in rows.items():
        fr = [d["roofline"]["frac"] for d in ds]
        summary[label] = {"frac_median": statistics.median(fr), "frac_all": fr,
                          "kernel_ms": [d["roofline"]["kernel_ms"] for d in ds],
                          "value": [d["value"] for d in ds], "workload": ds[0]["config"]["workload"]}
        print(f"== {label:16s} frac median {statistics.median(fr):.4f}  min {min(fr):.4f}  max {max(fr):.4f}")
    with open(os.path.join(ROOT, "gpurun_out", f"ab_{a.tag}.json"), "w") as f:
        json.dump(summary, f, indent=1)


if __name__ == "__main__":
    main()
