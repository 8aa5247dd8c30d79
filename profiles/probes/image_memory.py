"""Does the kind of memory the egocentric crops live in change the step?  (run on the GPU box)
    python profiles/probes/image_memory.py [--sizes 8192,65536]
Times the step's kernels (fused_vs_split.py times) with BCG_IMAGE_MEMORY = plain (torch tensor), vmm (the same virtual-
memory allocation without compression) and compressed (CU_MEM_ALLOCATION_COMP_GENERIC), then reads the crops back as a
GPU-resident consumer would (a sum over the image tensor) and checks that all three hold the same bytes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sizes = sys.argv[sys.argv.index("--sizes") + 1] if "--sizes" in sys.argv else "8192,65536"

if "--child" in sys.argv:
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "profiles", "probes"))
    import torch
    import fused_vs_split as F
    out = {}
    for n in [int(x) for x in sizes.split(",")]:
        env, actions = F.make_env(n)
        for k in range(40):
            env.step(actions[k % 16])
        torch.cuda.synchronize()
        img = env.ego_image
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v = img.view(-1)[: img.numel() // 8 * 8].view(torch.int64)
        for _ in range(3):
            v.sum()
        e0.record()
        for _ in range(10):
            v.sum()
        e1.record()
        torch.cuda.synchronize()
        out[n] = {"checksum": int(v.sum().item()), "nonzero": int(torch.count_nonzero(img).item()),
                  "read_ms": e0.elapsed_time(e1) / 10, "compressed": bool(getattr(env, "image_memory_compressed", False))}
        del env, actions, img, v
        torch.cuda.empty_cache()
    print(json.dumps(out))
    sys.exit(0)

for mode in ("plain", "vmm", "compressed"):
    env = dict(os.environ, BCG_IMAGE_MEMORY=mode)
    t = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "probes", "fused_vs_split.py"), "times", "--sizes", sizes],
                       env=env, capture_output=True, text=True)
    try:
        d = json.loads(t.stdout.strip().splitlines()[-1])["times"]
        for n, r in d.items():
            print("%-10s n=%-6s move %.4f reward %.4f ego %.4f step %.4f graph %.4f" % (
                mode, n, r["cr"], r["state_or_commit"], r["ego"], r["step_ms_plain"], r.get("step_ms_graph_same_actions", 0)), flush=True)
    except Exception:
        print(mode, "times failed", t.stdout[-300:], t.stderr[-800:], flush=True)
    c = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--sizes", sizes], env=env, capture_output=True, text=True)
    print("%-10s %s" % (mode, c.stdout.strip().splitlines()[-1] if c.stdout.strip() else "child failed: " + c.stderr[-800:]), flush=True)
