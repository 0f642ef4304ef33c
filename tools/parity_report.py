"""Print the kernel-vs-oracle error table (run on a GPU box):  python tools/parity_report.py [steps]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from taco_b200 import make_cfg
import parity_util as pu

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
report = {}
for name, kw, strict in [
    ("flip_strict", dict(task_mode="flip"), True),
    ("flip_fast", dict(task_mode="flip"), False),
    ("pos_strict", dict(task_mode="pos"), True),
    ("rotate_strict", dict(task_mode="rotate"), True),
    ("mix_dr_noise_strict", dict(task_mode="mix", domain_randomization=True, observation_noise=True, rotor_noise=True), True),
    ("mix_dr_noise_fast", dict(task_mode="mix", domain_randomization=True, observation_noise=True, rotor_noise=True), False),
]:
    cfg = make_cfg(num_envs=4096, **kw)
    gpu, ref = pu.make_pair(cfg, strict_fp=strict)
    res = pu.run_lockstep(gpu, ref, steps)
    worst_1 = res[0]
    allmis = {}
    for errs, mism, dmis, nfin in res:
        for k, v in mism.items():
            allmis[k] = allmis.get(k, 0) + v
        allmis["delayed_action"] = allmis.get("delayed_action", 0) + dmis
    report[name] = dict(step1_elementwise=ref.first_step_elementwise, step1=worst_1[0], step2=res[1][0], stepN=res[-1][0], mismatches=allmis, finite_last=res[-1][3],
                        stats_ref=ref.stats, stats_gpu=gpu.stats().cpu().tolist())
    print("==", name)
    print("  step 1 max rel err:", {k: "%.2e" % v for k, v in worst_1[0].items()})
    print("  step 1 element-wise rel err (floor 1e-2 / 1e-4):", {k: "%.2e" % v for k, v in ref.first_step_elementwise.items()})
    print("  step 2 max rel err:", {k: "%.2e" % v for k, v in res[1][0].items()})
    print("  step %d max rel err:" % steps, {k: "%.2e" % v for k, v in res[-1][0].items()})
    print("  integer/mask mismatches over all steps:", allmis, " finite envs at end:", res[-1][3])
    print("  stats ref:", ref.stats)
    print("  stats gpu:", gpu.stats().cpu().tolist())
    gpu.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(report, open("gpurun_out/parity_report.json", "w"), indent=1)
