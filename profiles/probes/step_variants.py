"""A/B of tuning builds of the state kernels on the bench workload (run on the GPU box):
    python profiles/probes/step_variants.py [--sizes 8192,65536]
Builds nothing: the variants are csrc/variants/libbcg_b200_<name>.so made by csrc/build.py build_variant (here, before
gpurun ships the tree) and are selected with BCG_B200_LIB; every build is timed by fused_vs_split.py times."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sizes = sys.argv[sys.argv.index("--sizes") + 1] if "--sizes" in sys.argv else "8192,65536"
only = sys.argv[sys.argv.index("--only") + 1].split(",") if "--only" in sys.argv else None
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, "bc_gym_planning_env_b200", "csrc", "variants", "*.so"))):
    name = "product" if lib is None else ("product, BCG_EGO_KERNEL=cta" if lib == "cta" else os.path.basename(lib)[len("libbcg_b200_"):-3])
    if only and name not in only:
        continue
    env = dict(os.environ)
    if lib == "cta":
        env["BCG_EGO_KERNEL"] = "cta"                 # round 1's CTA-per-env scatter kernel
    elif lib:
        env["BCG_B200_LIB"] = lib
    out = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "probes", "fused_vs_split.py"), "times", "--sizes", sizes],
                         env=env, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])["times"]
        for n, r in d.items():
            print("%-24s n=%-6s move %.4f reward %.4f ego %.4f step %.4f graph %.4f" % (
                name, n, r["cr"], r["state_or_commit"], r["ego"], r["step_ms_plain"], r.get("step_ms_graph_same_actions", 0)), flush=True)
    except Exception:
        print(name, "failed", out.stdout[-300:], out.stderr[-600:], flush=True)
