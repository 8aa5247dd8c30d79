"""A/B of the egocentric kernel variants on the bench workload (run on the GPU box):
    python profiles/probes/egt_exp.py
BCG_EGT_VARIANT 1 = one window per CTA, 5 CTAs/SM; 2 = two windows per CTA, 2 CTAs/SM.
BCG_EGT_DBG bit 0 skips the window loads, bit 1 the gather + stores, bit 2 the L2 prefetch (timing experiments only)."""
import json
import os
import subprocess
import sys

for staging, variant in (("tiles", 1), ("tma", 0)):
    for dbg in (0,):
        if staging == "tma" and dbg:
            continue
        env = dict(os.environ, BCG_EGT_DBG=str(dbg), BCG_EGT_VARIANT=str(variant))
        out = subprocess.run([sys.executable, "bench.py", "--steps", "50", "--warmup", "5", "--e2e-steps", "2", "--cpu-seconds", "0.1",
                              "--ego-staging", staging], env=env, capture_output=True, text=True)
        try:
            d = json.loads(out.stdout.strip().splitlines()[-1])
            print(staging, variant, "dbg", dbg, "ego_ms %.4f commit_ms %.4f step_ms %.4f" % (d["kernels_ms"].get("ego_tiles_kernel", d["kernels_ms"].get("ego_kernel")), d["kernels_ms"]["commit_kernel"], d["ms_per_step"]), flush=True)
        except Exception:
            print(staging, variant, "dbg", dbg, "failed", out.stderr[-800:], flush=True)
