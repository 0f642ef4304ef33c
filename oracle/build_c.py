"""Compile the oracle's C restatement (oracle/rigid_body.c) into oracle/librigid_body.so.  TEST INFRASTRUCTURE.
Called by __graft_entry__.build(); the .so is git-ignored and travels to the GPU box with the snapshot."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "rigid_body.c")
LIB = os.path.join(HERE, "librigid_body.so")


def build(force=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc") or shutil.which("cc")
    if not gcc:
        raise RuntimeError("gcc not found: cannot build the oracle's C integrator")
    subprocess.run([gcc, "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
