import sys, torch
sys.path.insert(0, '.')
import robot_gnn_b200 as R
from oracle import gnn_oracle as O
for (G, K, F, dt) in [(7, 5, 48, torch.float32), (7, 5, 64, torch.float32), (3, 5, 32, torch.bfloat16), (5, 3, 16, torch.float32), (5, 8, 64, torch.float32)]:
    sb = O.synthetic_batch(G, K, F, p=64, seed=1)
    z32 = torch.rand(G * K * (K - 1), 2 * F)
    g = R.robot_batch(G, K, device='cuda')
    h = sb['h'].reshape(G * K, F, 8, 8).cuda().to(dt)
    a = R.film_mean(g, h, z32.cuda().to(dt)).float().cpu()
    ref = O.film_mean(h.float().cpu().reshape(sb['h'].shape), z32.to(dt).float(), sb['src'], sb['dst']).reshape(a.shape)
    d = (a - ref).abs()
    nz = (d > 0).nonzero()
    print(G, K, F, dt, 'max', d.max().item(), 'count', len(nz), 'of', d.numel(), nz[:5].tolist())
