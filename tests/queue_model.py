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
    """Ring of (action, absolute end slot) runs; the slot clock is 10 * progress."""

    def __init__(self, delay):
        self.act = np.zeros((CAP, 4), dtype=np.float32)
        self.end = np.zeros(CAP, dtype=np.int64)
        self.head, self.n, self.len, self.progress, self.overflow = 0, 0, int(delay), 0, False
        if self.len > 0:
            self.act[0] = 0.0
            self.end[0] = self.len
            self.n = 1

    def step(self, action, T):
        clk = 10 * self.progress
        if self.len + T <= 100 and self.n < CAP:
            slot = (self.head + self.n) % CAP
            self.act[slot] = action
            self.end[slot] = clk + self.len + T
            self.n += 1
        else:
            self.overflow = True
        self.len += T
        cur, left = self.head, self.n
        dact = self.act[cur].copy() if left > 0 else np.zeros(4, dtype=np.float32)
        run_end = self.end[cur] if left > 0 else 0
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
        self.n, self.head = left, cur
        while self.n > 0 and run_end <= clk2:
            self.head = (self.head + 1) % CAP
            self.n -= 1
            if self.n > 0:
                run_end = self.end[self.head]
        self.len = max(self.len - 10, 0)
        return reads
