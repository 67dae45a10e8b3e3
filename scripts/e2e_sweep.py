"""dev tool: host-buffer call (pinned buffers, float64 rows) under different NBR_HOST_* settings."""
import ctypes, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    from nimrud_b200 import _lib, synth
    n = 10_000_000
    EDGES = (0.1, 0.2, 0.4, 0.8, 1.6); RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
    cloud = synth.urban_scene(n, seed=20, device="cuda")
    host_in = cloud.cpu().pin_memory()
    lib = _lib.lib()
    e_arr, e_p = _lib.f64_array(EDGES); r_arr, r_p = _lib.f64_array(RADII)
    for label, tdt, code in (("f64", torch.float64, _lib.F64), ("f32", torch.float32, _lib.F32)):
        host_out = torch.empty((n, 20), dtype=tdt).pin_memory()
        def step():
            _lib.check(lib.nbr_multiscale_features_host(ctypes.c_void_p(host_in.data_ptr()), _lib.F32, n, ctypes.c_void_p(host_in.data_ptr()),
                                                        _lib.F32, n, e_p, r_p, 5, ctypes.c_void_p(host_out.data_ptr()), code, 0, None))
        step(); step()
        t0 = time.perf_counter()
        for _ in range(5): step()
        dt = (time.perf_counter() - t0) / 5
        print("%s %s: %.2f ms  %.3f G" % (sys.argv[2], label, dt * 1e3, n * 5 / dt / 1e9), flush=True)
        del host_out
else:
    for env in ({}, {"NBR_HOST_PIECE_MB": "1"}, {"NBR_HOST_PIECE_MB": "2"}, {"NBR_HOST_PIECE_MB": "8"}, {"NBR_HOST_PIECE_MB": "16"},
                {"NBR_HOST_THREADS": "9"}, {"NBR_HOST_THREADS": "5"}, {"NBR_HOST_THREADS": "32"}, {"NBR_HOST_WIRE": "f64"},
                {"NBR_HOST_BATCH_ROWS": "65536"}, {"NBR_HOST_BATCH_ROWS": "131072"}):
        e = dict(os.environ); e.update(env)
        subprocess.run([sys.executable, os.path.abspath(__file__), "child", str(env)], env=e)
