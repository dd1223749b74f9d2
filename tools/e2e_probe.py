"""Where does the end-to-end step time go?  Times the graphed learner step with/without the per-step H2D
copy and with/without the per-step stats read-back."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import synthetic_host_batch, flags_ns, bind_to_gpu_numa_node
from torchbeast_b200 import learner, monobeast, optim

bind_to_gpu_numa_node(0)
T, B, A = 80, 32, 6
model = monobeast.AtariNet((4, 84, 84), A, True); actor = monobeast.AtariNet((4, 84, 84), A, True)
opt = optim.RMSprop(model, lr=0.00048, eps=0.01, alpha=0.99)
flags = flags_ns(T, B)
host = [synthetic_host_batch(T, B, A, i, True) for i in range(2)]
slots = [{k: v.cuda() for k, v in host[0].items()} for _ in range(2)]
state = model.initial_state(B)
gl = learner.GraphedLearner(flags, model, actor, opt, slots[0], state)
copy_stream = torch.cuda.Stream()

def run(h2d, stats, steps=20):
    ready = [torch.cuda.Event(), torch.cuda.Event()]; freed = [torch.cuda.Event(), torch.cuda.Event()]
    for e in freed: e.record()
    def stage(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(freed[s])
            if h2d:
                for k, v in host[i % 2].items(): slots[s][k].copy_(v, non_blocking=True)
            ready[s].record(copy_stream)
    torch.cuda.synchronize()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); stage(0)
    for i in range(steps):
        if i + 1 < steps: stage(i + 1)
        s = i % 2
        torch.cuda.current_stream().wait_event(ready[s])
        gl.step(slots[s], state, None); freed[s].record()
        if stats: gl.stats()
    z.record(); torch.cuda.synchronize()
    return a.elapsed_time(z) / steps

for h2d in (0, 1):
    for stats in (0, 1):
        run(h2d, stats, 5)
        print("h2d=%d stats=%d  %.3f ms/step" % (h2d, stats, run(h2d, stats)))
# H2D alone
torch.cuda.synchronize(); a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(10):
    for k, v in host[i % 2].items(): slots[i % 2][k].copy_(v, non_blocking=True)
z.record(); torch.cuda.synchronize(); print("h2d alone %.3f ms/step" % (a.elapsed_time(z) / 10))
