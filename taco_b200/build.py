"""Build libtaco_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m taco_b200.build            # or __graft_entry__.build()

Two translation units hold the fused step kernel: fpv_step_fast.cu (FMA contraction on) and
fpv_step_strict.cu (-fmad=false).  The result, taco_b200/lib/libtaco_b200.so, exposes only the
C ABI declared in include/taco_b200.h.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
_MB = os.environ.get("TACO_MIN_BLOCKS")          # tuning builds: libtaco_b200_mb<N>.so next to the default library
_XDEF = os.environ.get("TACO_XDEFS", "").split()  # extra -D flags for tuning builds (used with TACO_MIN_BLOCKS to get a separate .so)
LIB_PATH = os.path.join(LIB_DIR, "libtaco_b200.so" if not _MB else f"libtaco_b200_mb{_MB}.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--threads", "4"]
UNITS = [
    ("fpv_step_fast.cu", []),
    ("fpv_step_strict.cu", ["-fmad=false"]),
    ("taco_env.cu", []),
    ("taco_actor.cu", []),
    ("taco_critic.cu", []),
    ("taco_gae.cu", ["-fmad=false"]),
    ("taco_ppo.cu", []),
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build" if not _MB else f"build_mb{_MB}")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "taco_b200.h"))
    objs, procs = [], []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + ([f"-DTACO_MIN_BLOCKS={_MB}"] if _MB else []) + [f"-D{d}" for d in _XDEF] + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    if force or procs or _stale(LIB_PATH, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB_PATH] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
