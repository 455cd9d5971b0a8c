#!/usr/bin/env python
"""Same-box A/B of environment / library variants on the headline workload.

    python tools/ab_bench.py [--rounds R] [--steps K] [--workload W] [--events B] NAME=ENV1=V1,ENV2=V2 ...

Every variant runs ``bench.py --no-extra --no-cpu-baseline`` in its own process (the switches are read when the handle is created),
the variants interleaved R times; prints events/s, the SM clock under load and the per-category milliseconds of one evaluation.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

ap = argparse.ArgumentParser()
ap.add_argument("--rounds", type=int, default=2)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--workload", default="single_e")
ap.add_argument("--events", type=int, default=None)
ap.add_argument("--precision", default=None)
ap.add_argument("variants", nargs="+")
a = ap.parse_args()

for r in range(a.rounds):
    for v in a.variants:
        name, _, envs = v.partition("=")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, val = kv.partition("=")
            env[k] = val
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--no-extra", "--no-cpu-baseline", "--steps", str(a.steps), "--warmup", str(a.warmup), "--workload", a.workload]
        if a.events:
            cmd += ["--events", str(a.events)]
        if a.precision:
            cmd += ["--precision", a.precision]
        out = subprocess.run(cmd, env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(f"variant={name} FAILED rc={out.returncode}\n{out.stderr[-2000:]}", flush=True)
            continue
        d = json.loads(line[-1])
        cat = d.get("roofline", {}).get("per_category_ms") or d.get("per_category_ms") or {}
        print(f"variant={name} {d['value']:.0f} ev/s e2e {d['e2e']['value']:.0f} clk {d['clocks']['sm_mhz']:.0f} MHz {d['clocks'].get('power_w', 0):.0f} W  "
              + " ".join(f"{k}={x:.2f}" for k, x in cat.items()), flush=True)
