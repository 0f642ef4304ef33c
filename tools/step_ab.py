"""A/B of step-kernel builds on ONE box: for every library given, tools/step_windows.py in a fresh process (TACO_B200_LIB),
rounds interleaved so that box drift hits every variant alike.  Prints the steady-state ms/step (median of the windows after
step 100) per variant and workload.
usage: python tools/step_ab.py <rounds> <lib> [<lib> ...]   (lib = file name under taco_b200/lib)"""
import json, os, statistics, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rounds = int(sys.argv[1]); libs = sys.argv[2:]
work = [("flip", ["flip", "2097152", "10", "25"]), ("mixdr", ["mix", "2097152", "8", "25", "--dr"]), ("flip4096", ["flip", "4096", "6", "200"])]
if os.environ.get("AB_WORK"):
    work = [w for w in work if w[0] in os.environ["AB_WORK"].split(",")]
res = {}
for r in range(rounds):
    for lib in libs:
        env = dict(os.environ, TACO_B200_LIB=os.path.join(root, "taco_b200", "lib", lib))
        for name, a in work:
            out = subprocess.run([sys.executable, os.path.join(root, "tools", "step_windows.py")] + a, env=env, capture_output=True, text=True, timeout=600)
            try:
                w = json.loads(out.stdout.strip().splitlines()[-1])["windows"]
            except Exception:
                print("FAILED", lib, name, out.stderr[-500:]); continue
            late = [x["ms_per_step"] for x in w if x["steps"][0] >= 100] or [w[-1]["ms_per_step"]]
            res.setdefault((lib, name), []).append(statistics.median(late))
            if len(w) > 2:                           # the synchronised early steps (no resets yet, boost clocks): the driver's 20-step burst
                res.setdefault((lib, name + ":steps25-50"), []).append(w[1]["ms_per_step"])
for (lib, name), v in sorted(res.items(), key=lambda kv: (kv[0][1], kv[0][0])):
    print(json.dumps({"workload": name, "lib": lib, "ms_per_step": v}))
