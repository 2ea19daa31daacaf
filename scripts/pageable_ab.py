"""Host-array evaluation / fit with PAGEABLE arrays for several staging-pool sizes (one process per setting)."""
import os, sys, subprocess, time
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np
    import splpak_b200 as sp
    from splpak_b200 import synth
    nq = 100_000_000
    q = synth.queries_numpy(3, nq, seed=43)
    coef = np.random.default_rng(0).standard_normal(24 ** 3)
    ts = []
    for rep in range(4):
        t0 = time.perf_counter(); out, ie = sp.eval_batch(3, q, coef, [0.] * 3, [1.] * 3, [24] * 3); ts.append(time.perf_counter() - t0)
    x, y, w = synth.points_numpy(3, nq, start=0, seed=42)
    h = sp.FitHandle(3, [0.] * 3, [1.] * 3, [24] * 3, 1.0)
    tf = []
    for rep in range(3):
        h.reset(); t0 = time.perf_counter(); h.add_points(x, y, w); c, ie2 = h.compute(); tf.append(time.perf_counter() - t0)
    print(f"threads {os.environ.get('SPLPAK_B200_COPY_THREADS')} nproc {os.cpu_count()} eval 1e8 pageable: best {1e3*min(ts[1:]):.1f} ms  fit 1e8 pageable: best {1e3*min(tf[1:]):.1f} ms", flush=True)
else:
    for n in (4, 8, 12, 16, 24):
        env = dict(os.environ, SPLPAK_B200_COPY_THREADS=str(n))
        subprocess.run([sys.executable, __file__, "child"], env=env)
