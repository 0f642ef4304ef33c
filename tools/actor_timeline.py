"""Decode the CTA-0 timeline written by TACO_ACTOR_TIMELINE (csrc/taco_actor.cu): python tools/actor_timeline.py <file> [max_events]"""
import sys, numpy as np
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(3, -1)
mx = int(sys.argv[2]) if len(sys.argv) > 2 else 120
names = {0x10: "mma wait_a[0] begin", 0x11: "mma wait_a[1] begin", 0x20: "mma wait_a[0] end", 0x21: "mma wait_a[1] end", 0x30: "mma issued+commit[0]", 0x31: "mma issued+commit[1]",
         0x01: "epi obs staged+arrive", 0x02: "epi part0 D ready", 0x03: "epi part0 drained (early arrive)", 0x04: "epi part1 D ready", 0x05: "epi A written+arrive",
         0x06: "epi last-layer part D ready", 0x07: "epi tail done",
         0x08: "epi lstm step begin", 0x09: "epi x_{t+1} loaded", 0x0A: "epi lstm part0 math done", 0x0B: "epi lstm part1 math done", 0x0C: "epi x loads issued"}   # critic kernel (TACO_CRITIC_TIMELINE)
ev = []
for reg in range(3):
    for w in a[reg]:
        if w == 0: continue
        ev.append((int(w >> np.uint64(8)), reg, int(w & np.uint64(0xFF))))
ev.sort()
t0 = ev[0][0]
for t, reg, code in ev[:mx]:
    print(f"{t - t0:8d}  {'MMA ' if reg == 0 else 'EPI%d' % (reg - 1)}  {names.get(code, hex(code))}")
