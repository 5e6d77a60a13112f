import sys, subprocess, time, torch
sys.path.insert(0, '.')
sys.argv = ['x']
import importlib.util
spec = importlib.util.spec_from_file_location('fm', 'benchmarks/film_micro.py'); fm = importlib.util.module_from_spec(spec); spec.loader.exec_module(fm)
for name, fn in (('bwd', fm.bwd), ('fwd', fm.fwd), ('gemm', fm.gemm)):
    p = subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap', '--format=csv,noheader', '-lms', '100'], stdout=subprocess.PIPE, text=True)
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 3.0:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    p.terminate(); out = p.stdout.read().strip().split('\n')
    print(name, 'avg ms', e0.elapsed_time(e1) / n, 'samples:', out[3:28:4])
