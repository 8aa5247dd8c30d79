"""CPU: the C-ABI shared library builds, loads, exports every symbol include/bcg_b200.h declares and
agrees with the ctypes mirrors.  No compute call needs a GPU here; the ones that do must fail loudly."""
import ctypes as C
import os
import re

import pytest

from bc_gym_planning_env_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bcg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bcg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = nat.lib()
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), "libbcg_b200.so does not export %s" % name
    assert sorted(nat.SYMBOLS) == declared, "ctypes binding and header disagree"
    assert lib.bcg_abi_version() == 11


def test_struct_mirrors_match_library():
    lib = nat.lib()
    for which, struct in enumerate(nat._STRUCTS):
        assert lib.bcg_sizeof(which) == C.sizeof(struct), struct.__name__
    assert lib.bcg_sizeof(99) == -1


def test_state_layout_follows_delays():
    L0 = nat.state_layout(nat.BcgParams())
    assert (L0.n_frows, L0.n_irows) == (nat.F_FIXED, nat.I_FIXED)
    L = nat.state_layout(nat.BcgParams(delay_control=2, delay_pose=1, delay_state=3))
    assert L.ring_control == nat.F_FIXED and L.ring_pose == L.ring_control + 4
    assert L.ring_state == L.ring_pose + 3 and L.n_frows == L.ring_state + 21
    with pytest.raises(nat.BcgError):
        nat.state_layout(nat.BcgParams(delay_pose=-1))


def test_errors_are_reported_not_thrown():
    lib = nat.lib()
    assert lib.bcg_step(None, None, None, 0, 0, None, None) < 0
    assert "null" in nat.last_error()
    assert lib.bcg_world_to_pixel(None, 4, 0.0, 0.0, 0.05, None, None) < 0
    p, b = nat.BcgParams(), nat.BcgBatch()
    assert lib.bcg_step(C.byref(p), C.byref(b), None, 0, 0, None, None) < 0
    assert "n_envs" in nat.last_error()


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(nat.BcgError):
        nat.require_cuda()
    from bc_gym_planning_env_b200.envs.base.params import EnvParams
    from bc_gym_planning_env_b200.envs.synth_turn_env import random_aisle_pool
    from bc_gym_planning_env_b200.vec_env import VecPlanEnv
    costmaps, paths = random_aisle_pool(1, 0)
    with pytest.raises(nat.BcgError):
        VecPlanEnv(costmaps, paths, EnvParams())


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "bc_gym_planning_env_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
