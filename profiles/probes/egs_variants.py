"""A/B of kernel shapes (ego_sparse_kernel CTA shape, block sizes of the state kernels) on the bench workload (run on the GPU box):
    python profiles/probes/egs_variants.py
Builds nothing: the variants are csrc/variants/libbcg_b200_<name>.so made by csrc/build.py build_variant and are
selected with BCG_B200_LIB."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "bc_gym_planning_env_b200", "csrc", "variants", "*.so"))):
    env = dict(os.environ)
    if lib:
        env["BCG_B200_LIB"] = lib
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "60", "--warmup", "5", "--e2e-steps", "2",
                          "--cpu-seconds", "0.1", "--gen-envs", "0"], env=env, capture_output=True, text=True)
    name = os.path.basename(lib) if lib else "product build"
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        k = d["kernels_ms"]
        print("%-32s kin %.4f collide %.4f commit %.4f ego %.4f step_ms %.4f value %.4g dense-fallback envs %s" % (
            name, k["kin_kernel"], k["collide_reward_kernel"], k["commit_kernel"], k["ego_sparse_kernel"], d["ms_per_step"],
            d["value"], d["ego_dense_fallback_envs_last_step"]), flush=True)
    except Exception:
        print(name, "failed", out.stderr[-600:], flush=True)
