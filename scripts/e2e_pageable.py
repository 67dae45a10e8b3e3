"""dev tool: where the numpy shim loses against pinned buffers -- host-buffer call with pinned / pageable input and output."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import _lib, synth, multiscale
n = 10_000_000
EDGES = (0.1, 0.2, 0.4, 0.8, 1.6); RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
cloud = synth.urban_scene(n, seed=20, device="cuda")
pin_in = cloud.cpu().pin_memory()
page_in = pin_in.numpy().copy()
lib = _lib.lib()
e_arr, e_p = _lib.f64_array(EDGES); r_arr, r_p = _lib.f64_array(RADII)
pin_out = torch.empty((n, 20), dtype=torch.float64).pin_memory()
page_out = np.empty((n, 20), dtype=np.float64); page_out[...] = 0
def call(inp, outp):
    _lib.check(lib.nbr_multiscale_features_host(ctypes.c_void_p(inp), _lib.F32, n, ctypes.c_void_p(inp), _lib.F32, n, e_p, r_p, 5,
                                                ctypes.c_void_p(outp), _lib.F64, 0, None))
def timed(label, fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    dt = (time.perf_counter() - t0) / reps
    print("%-46s %.2f ms  %.3f G" % (label, dt * 1e3, n * 5 / dt / 1e9), flush=True)
timed("pinned in, pinned out", lambda: call(pin_in.data_ptr(), pin_out.data_ptr()))
timed("pageable in, pinned out", lambda: call(page_in.ctypes.data, pin_out.data_ptr()))
timed("pinned in, pageable out (touched)", lambda: call(pin_in.data_ptr(), page_out.ctypes.data))
timed("pageable in, pageable out (touched)", lambda: call(page_in.ctypes.data, page_out.ctypes.data))
timed("process_single_core(ndarray), result dropped", lambda: multiscale.process_single_core(page_in, page_in, EDGES, RADII))
keep = [None]
def kept():
    keep[0] = multiscale.process_single_core(page_in, page_in, EDGES, RADII)
timed("process_single_core(ndarray), result kept", kept)
