"""Python model of the kernel's run-length pending-action queue (csrc/fpv_step_kernel.cuh) and of the
reference's dense (4,100) delay buffer (fpv_asymmetry.py:189,326-331,366,378-380), for property tests."""
import numpy as np

CAP = 16


class DenseDelay:
    def __init__(self, delay):
        self.buf = np.zeros((4, 100), dtype=np.float32)
        self.len = int(delay)
        self.overflow = False

    def step(self, action, T):
        lo, hi = self.len, self.len + T
        if hi > 100:
            self.overflow = True
        self.buf[:, max(lo, 0):min(hi, 100)] = np.asarray(action, dtype=np.float32)[:, None]
        self.len += T
        reads = []
        for k in range(10):
            idx = min(self.len - 1, k)          # negative wraps like python/torch
            reads.append(self.buf[:, idx].copy())
        self.buf[:, 0:-10] = self.buf[:, 10:].copy()
        self.len = max(self.len - 10, 0)
        return reads


class RunQueue:
    """Ring of (action, absolute end slot) runs; the slot clock is 10 * progress.  The ring position of a run is the
    RL step index of the LAUNCH that appended it (t & 15, identical for every env of a launch); an env's live runs are
    the last n positions, so no head pointer is stored."""

    def __init__(self, delay, t0=0):
        self.act = np.zeros((CAP, 4), dtype=np.float32)
        self.end = np.zeros(CAP, dtype=np.int64)
        self.t = int(t0)                     # global step index at which the env was (lazily) reset
        self.n, self.len, self.progress, self.overflow = 0, int(delay), 0, False
        if self.len > 0:
            z = (self.t - 1) % CAP
            self.act[z] = 0.0
            self.end[z] = self.len
            self.n = 1

    def step(self, action, T):
        clk = 10 * self.progress
        ts = self.t % CAP
        if self.len + T > 100 or self.n >= CAP:
            self.overflow = True
            if self.n >= CAP:
                self.n = CAP - 1
        self.act[ts] = action
        self.end[ts] = clk + self.len + T
        self.n += 1
        self.len += T
        cur, left = (ts + 1 - self.n) % CAP, self.n
        dact, run_end = self.act[cur].copy(), self.end[cur]
        reads = []
        for k in range(10):
            slot_abs = clk + min(self.len - 1, k)
            while slot_abs >= run_end and left > 1:
                cur = (cur + 1) % CAP
                left -= 1
                dact, run_end = self.act[cur].copy(), self.end[cur]
            reads.append(dact.copy())
        self.progress += 1
        clk2 = clk + 10
        self.n = left
        while self.n > 0 and run_end <= clk2:
            cur = (cur + 1) % CAP
            self.n -= 1
            if self.n > 0:
                run_end = self.end[cur]
        self.len = max(self.len - 10, 0)
        self.t += 1
        return reads
