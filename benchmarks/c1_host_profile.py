import sys, torch, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/benchmarks')
import configs
dev=torch.device('cuda:0')
step=configs.c1(dev)
for _ in range(5): step()
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(50): step()
t1=time.perf_counter()
torch.cuda.synchronize()
t2=time.perf_counter()
print('host enqueue per step ms', (t1-t0)/50*1e3, 'incl drain', (t2-t0)/50*1e3)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
  for _ in range(10): step()
  torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=35, max_name_column_width=60))
