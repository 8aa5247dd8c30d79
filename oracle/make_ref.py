"""TEST INFRASTRUCTURE ONLY -- installs the UNMODIFIED upstream reference into oracle/_ref/ (git-ignored).

    python -m oracle.make_ref [--force]

The reference (braincorp/bc-gym-planning-env, MIT) is pure Python: "building" it is a `pip install --no-deps --target`
of the tree under /root/reference (from a scratch copy, because setuptools writes egg-info next to setup.py and the
mount is read-only).  oracle/_ref/ is listed in .gitignore, not in .gpurunignore, so the installed package travels to
the GPU box with the snapshot -- there it is what `bench.py --impl reference`, the `cpu_baseline` leg and
tests/test_gpu_live_reference.py run.  Nothing of it is committed, and the product package never imports it
(oracle/ref_loader.py is the only way in).
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
TARGET = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("BCG_REFERENCE_SOURCE", "/root/reference")


def installed():
    return os.path.isfile(os.path.join(TARGET, "bc_gym_planning_env", "envs", "base", "env.py"))


def make_ref(force=False):
    """Returns the install directory, or None when there is no reference tree to install from."""
    if installed() and not force:
        return TARGET
    if not os.path.isfile(os.path.join(SOURCE, "setup.py")):
        return TARGET if installed() else None
    scratch = tempfile.mkdtemp(prefix="bcg_ref_")
    try:
        src = os.path.join(scratch, "src")
        shutil.copytree(SOURCE, src, ignore=shutil.ignore_patterns(".git", "__pycache__", "img"))
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        cmd = [sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
               "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=scratch)
        if res.returncode != 0:
            raise RuntimeError("pip install of the reference failed:\n%s\n%s" % (res.stdout[-2000:], res.stderr[-2000:]))
        lic = os.path.join(SOURCE, "LICENSE")
        if os.path.isfile(lic):
            shutil.copy(lic, os.path.join(TARGET, "LICENSE.bc_gym_planning_env"))
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return TARGET


if __name__ == "__main__":
    print(make_ref(force="--force" in sys.argv))
