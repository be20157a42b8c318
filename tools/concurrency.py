import sys, time, threading
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import numpy as np, torch
from ug_stereomatcher_b200 import capi, synth
L, R, _, _ = synth.make_pair(4928, 3264, 0)
dev = torch.device("cuda", 0)
dL = torch.from_numpy(L).to(dev); dR = torch.from_numpy(R).to(dev)
def run(nh, steps=6):
    hs = [capi.Handle() for _ in range(nh)]
    outs = [[torch.empty((3264, 4928), dtype=torch.float32, device=dev) for _ in range(3)] for _ in range(nh)]
    def work(k, n):
        for _ in range(n):
            hs[k].match_full_device(dL.data_ptr(), dR.data_ptr(), 3264, 4928, 4928 * 3, *[o.data_ptr() for o in outs[k]])
    for k in range(nh): work(k, 2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(k, steps)) for k in range(nh)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("handles=%d: %.2f ms per pair (throughput), %.1f pairs/s" % (nh, dt * 1e3 / (nh * steps), nh * steps / dt))
    t = hs[0].timing(); print("   last device ms %.2f level0 %.2f k1_l0 %.3f k2_l0 %.3f" % (t.total_ms, t.level_ms[0], t.k1_ms[0], t.k2_ms[0]))
    [h.close() for h in hs]
for nh in (1, 2, 3): run(nh)
