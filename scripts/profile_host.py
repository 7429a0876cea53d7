"""Host-side cost of issuing one schedule step: cProfile of the main thread (forward + engine wait) and the
forward / backward split.  The step is device-bound only while the host stays ahead (~6 ms vs ~8.5 ms)."""
import cProfile, io, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ao_b200 import scenes
from ao_b200 import schedule as S

dev = torch.device("cuda", 0)
coord_np, _, off_np = scenes.s3dis_batch(4, 80000)
coord, offset = torch.from_numpy(coord_np).to(dev), torch.from_numpy(off_np).to(dev)
sched = S.PointOpsSchedule(S.ScheduleConfig.s3dis(), device=dev, seed=0)
for _ in range(3):
    sched.step(coord, offset)
torch.cuda.synchronize()
# forward/backward split: wrap autograd.grad
orig = torch.autograd.grad
acc = {"bwd": 0.0}
def timed_grad(*a, **k):
    t = time.perf_counter(); r = orig(*a, **k); acc["bwd"] += time.perf_counter() - t; return r
torch.autograd.grad = timed_grad
t0 = time.perf_counter()
for _ in range(5):
    sched.step(coord, offset)
tot = time.perf_counter() - t0
torch.cuda.synchronize()
print("host enqueue per step: total %.2f ms, backward (autograd.grad) %.2f ms, forward %.2f ms" % (tot / 5 * 1e3, acc["bwd"] / 5 * 1e3, (tot - acc["bwd"]) / 5 * 1e3))
torch.autograd.grad = orig
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    sched.step(coord, offset)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(32)
print(s.getvalue()[:6000])
