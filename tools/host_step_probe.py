"""Times taco_env_step_host variants (mapped / copy mode, with and without result buffers) against the device-resident step.
usage: python tools/host_step_probe.py [task] [num_envs] [steps]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import taco_b200  # noqa: E402


def main():
    task = sys.argv[1] if len(sys.argv) > 1 else "flip"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2097152
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    env = taco_b200.FpvVecTask(taco_b200.make_cfg(task, n), "cuda:0", "cuda:0", -1, True, seed=1)
    acts = [env.random_actions(t) for t in range(4)]
    h_act = [a.cpu().pin_memory() for a in acts]
    h_rew = torch.empty(n).pin_memory(); h_reset = torch.empty(n, dtype=torch.int64).pin_memory()
    h_tout = torch.empty(n, dtype=torch.uint8).pin_memory()
    for t in range(100):                                   # desynchronise the episodes first
        env.step(acts[t % 4])

    def timed(fn):
        for t in range(3):
            fn(t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for t in range(k):
            fn(t)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    out = {"task": task, "num_envs": n, "steps": k}
    out["device_ms"] = timed(lambda t: env.step(acts[t % 4]))
    for mode in ("mapped", "copy"):
        os.environ["TACO_HOST_MODE"] = mode
        out[mode + "_full_ms"] = timed(lambda t: env.step_host(h_act[t % 4], h_rew, h_reset, h_tout))
        out[mode + "_in_only_ms"] = timed(lambda t: env.step_host(h_act[t % 4]))
        out[mode + "_in_rew_ms"] = timed(lambda t: env.step_host(h_act[t % 4], h_rew))
        out[mode + "_in_rew_reset_ms"] = timed(lambda t: env.step_host(h_act[t % 4], h_rew, h_reset))
    for key in list(out):
        if key.endswith("_ms"):
            out[key.replace("_ms", "_env_steps_per_s")] = n / (out[key] * 1e-3)
    print(json.dumps(out))
    env.close()


if __name__ == "__main__":
    main()
