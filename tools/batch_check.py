import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbt_b200.generator import GeneratorJ
def timeit(fn, warm=3, reps=6):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/reps
torch.manual_seed(0)
g = GeneratorJ(input_channels=3, use_bias=True).cuda().eval()
for n in (1,2,4):
    x = torch.rand(n,3,1080,1920,device="cuda")*2-1
    with torch.no_grad(): ms = timeit(lambda: g(x))
    print(f"batch {n}: {ms:.2f} ms per pass -> {ms/n:.2f} ms/frame, {n*1e3/ms:.1f} fps", flush=True)
    del x; g._engine._ws.clear(); torch.cuda.empty_cache()
