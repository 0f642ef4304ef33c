"""Time the actor kernels alone: python tools/actor_bench.py [envs] [hidden,comma] [iters]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, taco_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
hidden = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "256,256,256").split(",")]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
sizes = [26] + hidden + [4]
gen = torch.Generator().manual_seed(1)
ws = [torch.randn(sizes[l + 1], sizes[l], generator=gen) * (1.0 / sizes[l] ** 0.5) for l in range(len(sizes) - 1)]
bs = [torch.randn(sizes[l + 1], generator=gen) * 0.1 for l in range(len(sizes) - 1)]
a = taco_b200.ActorMLP(26, hidden, 4)
a.load(ws, bs)
obs = torch.randn(n, 26, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
out = torch.empty(n, 4, device="cuda")
def t(fn, it):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
ms_tc = t(lambda: a.forward(obs, tensor_cores=True, out=out), iters)
flops = 2.0 * sum(sizes[i] * sizes[i + 1] for i in range(len(sizes) - 1)) * n
ref = a.forward(obs[:4096], tensor_cores=False)
got = a.forward(obs[:4096], tensor_cores=True)
print(json.dumps({"envs": n, "sizes": sizes, "tc_ms": ms_tc, "tflops": flops / ms_tc / 1e9, "max_abs_diff_vs_fp32_kernel": float((ref - got).abs().max())}))
