import sys, torch
sys.path.insert(0, '.')
import robot_gnn_b200 as R
F, L = R.functional, R._lib
m, n, k = 20480, 2560, 1280
x = torch.randn((m, k), device='cuda'); w = torch.randn((n, k), device='cuda'); y = torch.empty((m, n), device='cuda')
def run(): F._gemm(x, (0, 0, k, 1), w, (0, 0, 1, k), y, (0, n, 1), m, n, k)
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('linear %.3f ms  %.1f TF/s' % (ms, 2.0 * m * n * k / ms / 1e9))
