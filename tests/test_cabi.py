"""The C-ABI library loads and exports every symbol include/taco_b200.h declares.  CPU only:
no compute entry point is called (there is no GPU here); argument validation that happens
before any CUDA call is exercised."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from taco_b200 import build, _capi
    build.build()
    return _capi.lib()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "taco_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(taco_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = _declared_functions()
    for must in ("taco_env_create", "taco_env_destroy", "taco_env_buffers", "taco_env_step", "taco_env_step_host",
                 "taco_env_reset_all", "taco_env_set_difficulty", "taco_env_set_seed", "taco_env_stats",
                 "taco_actor_forward", "taco_last_error"):
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/taco_b200.h but not exported"


def test_abi_version_and_struct_layout(lib):
    from taco_b200 import _capi
    assert lib.taco_abi_version() == _capi.ABI_VERSION == 1
    # TacoCfg: 2 x i32, 2 x i64, 8 x i32/u32, 4 x f32, u64 -> 80 bytes with natural alignment
    assert C.sizeof(_capi.TacoCfg) == 80
    assert _capi.TacoCfg.env_offset.offset == 8 and _capi.TacoCfg.seed.offset == 72


def test_null_arguments_are_rejected_without_cuda(lib):
    rc = lib.taco_env_create(None, 0, None)
    assert rc == -1 and b"null" in lib.taco_last_error()
    assert lib.taco_env_step(None, None, None) == -1
    assert lib.taco_env_destroy(None) == 0


def test_bad_config_is_rejected_before_touching_the_device(lib):
    from taco_b200 import _capi
    cfg = _capi.TacoCfg(abi_version=99, num_envs=4)
    h = C.c_void_p()
    assert lib.taco_env_create(C.byref(cfg), 0, C.byref(h)) == -1
    assert b"ABI" in lib.taco_last_error()
    cfg = _capi.TacoCfg(abi_version=1, num_envs=0)
    assert lib.taco_env_create(C.byref(cfg), 0, C.byref(h)) == -1


def test_product_path_fails_loudly_without_cuda():
    import torch
    import taco_b200
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        taco_b200.FpvFlip(taco_b200.make_cfg("flip", 16))


def test_product_does_not_import_oracle():
    import subprocess, sys
    code = "import sys; import taco_b200, taco_b200._capi, taco_b200.fpv_vec_task; print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))"
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, check=True).stdout.strip()
    assert out == "False"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "taco_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                # comments may CITE oracle files (oracle/..., oracle.philox ...); nothing may import, load or execute them
                for pat in (r"^\s*(from|import)\s+oracle\b", r"^\s*from\s+\.+\s*oracle\b", r"import_module\(\s*['\"]oracle",
                            r"__import__\(\s*['\"]oracle", r"librigid_body", r"dlopen\([^)]*oracle", r"CDLL\([^)]*oracle",
                            r"#include\s*[<\"][^>\"]*oracle"):
                    assert not re.search(pat, src, flags=re.M), (f, pat)


# ---------------------------------------------------------------------------------------------- a plain-C consumer
def _build_consumer(tmp_path):
    import shutil
    import subprocess
    gcc = shutil.which("gcc") or shutil.which("cc")
    if gcc is None:
        pytest.skip("no C compiler")
    lib_dir = os.path.join(ROOT, "taco_b200", "lib")
    exe = str(tmp_path / "consumer")
    cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi", "consumer.c"),
           "-L", lib_dir, "-ltaco_b200", f"-Wl,-rpath,{lib_dir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_plain_c_and_a_c_program_links(lib, tmp_path):
    """include/taco_b200.h compiles as C99 with -Wall -Werror and a consumer without CUDA headers or torch links against the .so."""
    assert os.path.exists(_build_consumer(tmp_path))


@pytest.mark.gpu
def test_c_consumer_reproduces_the_python_binding(tmp_path):
    """tests/cabi/consumer.c steps a flip env through taco_env_step_host with malloc'ed buffers; FpvVecTask with the same
    configuration, seed and action pattern must produce the same rewards, resets and time-outs."""
    import re
    import subprocess
    import torch
    import taco_b200
    n, steps = 1000, 12
    exe = _build_consumer(tmp_path)
    r = subprocess.run([exe, str(n), str(steps)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    m = re.search(r"sum_rew=(\S+) n_reset=(\d+) n_tout=(\d+) stats_sum_rew=(\S+) stats_n_done=(\d+) stats_env_steps=(\d+)", r.stdout)
    assert m, r.stdout
    cfg = taco_b200.make_cfg("flip", n, random_voltage=False, random_rotor_speed=False, **{"env.clipActions": 1.0})
    env = taco_b200.FpvVecTask(cfg, "cuda:0", "cuda:0", -1, True, seed=5)
    i = torch.arange(n)                                  # the pattern is built on the CPU: torch's CUDA division by a scalar multiplies by the reciprocal
    sum_rew, n_reset, n_tout = 0.0, 0, 0
    for t in range(steps):
        act = torch.stack([((i * 7 + t * 3) % 17).float() / 16.0 - 0.5, ((i * 5 + t) % 13).float() / 24.0 - 0.25,
                           ((i * 3 + t * 2) % 11).float() / 20.0 - 0.25, ((i + t * 5) % 7).float() / 12.0 - 0.25], dim=1).contiguous().cuda()
        _, rew, reset, extras = env.step(act)
        sum_rew += float(rew.double().sum()); n_reset += int((reset != 0).sum()); n_tout += int(extras["time_outs"].sum())
    assert int(m.group(2)) == n_reset and int(m.group(3)) == n_tout
    assert float(m.group(1)) == pytest.approx(sum_rew, rel=1e-10)     # identical float32 rewards, summed in double in two orders
    stats = env.stats().cpu()
    assert float(m.group(4)) == pytest.approx(float(stats[0]), rel=1e-6) and int(m.group(5)) == int(stats[1]) and int(m.group(6)) == n * steps
    env.close()
