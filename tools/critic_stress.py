import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, taco_b200
n=262144; hid=64; mlp=[256,256,256]; T=5; IN=26
gen = torch.Generator().manual_seed(1)
lstm = [(torch.randn(4*hid, IN, generator=gen)*0.2, torch.randn(4*hid, hid, generator=gen)*0.15, torch.randn(4*hid, generator=gen)*0.1, torch.randn(4*hid, generator=gen)*0.1)]
sizes=[hid]+mlp+[1]
ws=[torch.randn(sizes[l+1], sizes[l], generator=gen)*(1.0/sizes[l]**0.5) for l in range(len(sizes)-1)]
bs=[torch.randn(sizes[l+1], generator=gen)*0.1 for l in range(len(sizes)-1)]
c=taco_b200.CriticLSTM(IN,T,hid,mlp); c.load(lstm,ws,bs)
states=torch.randn(n,T,IN,device="cuda"); out=torch.empty(n,1,device="cuda")
for _ in range(5): c.forward(states,tensor_cores=True,out=out)
torch.cuda.synchronize()
N=1500
evs=[(torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)) for _ in range(N)]
for e0,e1 in evs:
    e0.record(); c.forward(states,tensor_cores=True,out=out); e1.record()
torch.cuda.synchronize()
ts=sorted(e0.elapsed_time(e1) for e0,e1 in evs)
print(json.dumps({"calls":N,"min_ms":ts[0],"median_ms":ts[N//2],"p99_ms":ts[int(N*0.99)],"max_ms":ts[-1]}))
