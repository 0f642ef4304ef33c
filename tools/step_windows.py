"""ms/step of the fused step in windows of W steps (how the cost evolves as episodes desynchronise).
usage: python tools/step_windows.py [task] [envs] [windows] [W] [--dr] [--fast]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, taco_b200
args = [a for a in sys.argv[1:] if not a.startswith("--")]
task = args[0] if len(args) > 0 else "flip"
n = int(args[1]) if len(args) > 1 else 2 * 1024 * 1024
nw = int(args[2]) if len(args) > 2 else 16
W = int(args[3]) if len(args) > 3 else 25
env = taco_b200.FpvVecTask(taco_b200.make_cfg(task, n, domain_randomization="--dr" in sys.argv), "cuda:0", "cuda:0", -1, True, seed=0x7AC0, strict_fp="--fast" not in sys.argv)
acts = [env.random_actions(t) for t in range(4)]
out = []
k = 0
for w in range(nw):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(W):
        env.step(acts[k % 4]); k += 1
    e1.record(); torch.cuda.synchronize()
    st = env.stats().cpu().tolist()
    out.append({"steps": [k - W, k], "ms_per_step": round(e0.elapsed_time(e1) / W, 4), "done_rate": round(st[1] / max(st[7], 1), 5)})
print(json.dumps({"task": task, "envs": n, "windows": out}))
