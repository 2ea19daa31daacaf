"""Regular-grid evaluation (splpak_b200_eval_grid_device) against the point-wise kernel on the same raster-ordered
points: cfg3 table (24^3 nodes), n^3 output grid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import splpak_b200 as sp
nodes = [24, 24, 24]
coef = torch.randn(24 ** 3, dtype=torch.float64, device="cuda")
print("| grid | points | eval_grid ms | Gpoints/s | GB/s written | of HBM 6534 | point-wise (raster order) ms | speed-up |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for n in (256, 512, 1000):
    ax = torch.linspace(0.0, 1.0, n, dtype=torch.float64, device="cuda")
    d_axes = torch.cat([ax, ax, ax])
    nq = n ** 3
    out = torch.empty(nq, dtype=torch.float64, device="cuda")
    def timeit(fn, reps=4):
        best = 1e30
        for r in range(reps + 1):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if r: best = min(best, e0.elapsed_time(e1))
        return best
    tg = timeit(lambda: sp.eval_grid_device(3, d_axes, [n, n, n], coef, [0] * 3, [1] * 3, nodes, out, stream=torch.cuda.current_stream()))
    chk = out[:: max(1, nq // 1000)].clone()
    # the same points, point-wise, dimension 1 fastest
    idx = torch.arange(nq, device="cuda")
    q = torch.stack([ax[idx % n], ax[(idx // n) % n], ax[idx // (n * n)]], dim=1).contiguous()
    del idx
    tp = timeit(lambda: sp.eval_batch_device(3, q, 3, nq, coef, [0] * 3, [1] * 3, nodes, out, stream=torch.cuda.current_stream()))
    err = float((out[:: max(1, nq // 1000)] - chk).abs().max())
    print(f"| {n}^3 | {nq:.2e} | {tg:.3f} | {nq / tg / 1e6:.1f} | {nq * 8 / tg / 1e6:.0f} | {100 * nq * 8 / tg / 1e6 / 6534.5:.1f} % | {tp:.3f} | {tp / tg:.1f}x |  (max |grid - pointwise| on a sample: {err:.1e})", flush=True)
    del q
