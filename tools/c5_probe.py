"""Development probe: where one C5 training step goes (torch.profiler kernel table + wall vs device time)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
captured = {}
def timed(fn, steps):
    captured["fn"] = fn
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
print(bench.train_leg(0, 1, dev, timed))
fn = captured["fn"]
t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0):.1f} ms, until done {1e3*(t2-t0):.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fn(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
