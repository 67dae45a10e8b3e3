#!/usr/bin/env python
"""
bench.py -- points*scales/sec of the multiscale eigenfeature path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--points P]

One "step" = one pass of the whole hot path over the workload: bounding box, one lattice index per
scale, fused radius query + covariance + eigensolve + feature emission for every query at every scale.
Workload at N=1 = BASELINE.json configs[1]: 10M-point synthetic urban scene, 5 radius scales
(edge 0.1..1.6, r = 3e), query cloud == search cloud.  N>1: one such tile per GPU (weak scaling),
tiles side by side, lattices anchored on the all-reduced bounding box, halo exchange over NCCL.

value      whole-job points*scales/s with the cloud resident in HBM (CUDA events, max over ranks)
e2e        same through the reference-facing call with HOST buffers (pinned), H2D + D2H inside
roofline   the dominant kernel (fused feature kernel) against the measured HBM peak
cpu_baseline / --impl reference: the CPU port of the reference (oracle/nimrud_oracle.py, the same
           per-neighborhood numpy calls the reference makes) on the box's host cores, bounded sample
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

EDGES = (0.1, 0.2, 0.4, 0.8, 1.6)
RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
METRIC = "points_scales_per_sec"
UNIT = "points*scales/s"
FALLBACK_HBM_GBS = 6650.0


# --------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# --------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# --------------------------------------------------------------------------------------------------
def cpu_sample_cloud(points_target):
    """a spatially contiguous tile of the config-2 scene plus its halo, generated on the CPU."""
    import torch
    from nimrud_b200 import synth
    # a 2M-point scene at the config's density (about 224 m x 224 m) is enough to cut tiles from
    scene = synth.urban_scene(2_000_000, seed=20, device="cpu").numpy()
    halo = max(RADII) + max(EDGES) / 2
    side = math.sqrt(points_target / 40.0)
    lo = np.array([60.0, 60.0])
    inside = np.all((scene[:, :2] >= lo) & (scene[:, :2] < lo + side), axis=1)
    near = np.all((scene[:, :2] >= lo - halo) & (scene[:, :2] < lo + side + halo), axis=1)
    return scene[inside].astype(np.float64), scene[near].astype(np.float64)


def run_cpu_reference(steps, warmup, workers):
    from oracle import nimrud_oracle as O
    per_worker = 6000                                   # queries per worker per step (~10 s of CPU work)
    query, search = cpu_sample_cloud(per_worker * workers)
    query = query[:per_worker * workers]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        out = O.process_parallel(query, search, EDGES, RADII, workers)
        dt = time.perf_counter() - t0
        assert out.shape == (len(query), 4 * len(RADII))
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    value = len(query) * len(RADII) * len(times) / total
    sample = "%d contiguous queries (tile of the 10M-scene generator) x %d scales against tile+halo (%d points), " \
             "%d processes" % (len(query), len(RADII), len(search), workers)
    return value, 1e3 * total / len(times), sample


def reference_arm(args, rank):
    if rank != 0:
        return
    workers = max(1, min(os.cpu_count() or 1, 64))
    steps = max(1, min(args.steps, 3))
    value, ms, sample = run_cpu_reference(steps, min(args.warmup, 1), workers)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.points, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(points, gpus):
    return {"workload": "BASELINE configs[1]: %dM-point synthetic urban scene per GPU, 5 radius scales "
                        "(edge 0.1/0.2/0.4/0.8/1.6 m, radius 3*edge), query == search" % (points // 1_000_000),
            "points_per_gpu": points, "scales": len(RADII), "edges": list(EDGES), "radii": list(RADII),
            "l2_policy": "inputs+outputs per step (920 MB) exceed the 126 MB L2; no explicit flush",
            "parallelism": "spatial tiles, 1 per GPU" if gpus > 1 else "single GPU"}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def seam_check(nd, multiscale, dist, cloud, out, edges, radii, rank, world):
    """rank 0 recomputes, on ONE GPU and against the union of all tiles, the rows of its queries that lie within the
    halo width of its tile's border, and compares them bit for bit with what the tile path produced."""
    import torch
    n = cloud.shape[0]
    union = torch.empty((world * n, 3), dtype=cloud.dtype, device=cloud.device)
    dist.all_gather_into_tensor(union, cloud)
    result = None
    if rank == 0:
        h = nd.halo_width(edges, radii)
        # the queries that can need foreign points: those inside another tile's box grown by h
        near_mask = torch.zeros(n, dtype=torch.bool, device=cloud.device)
        for r in range(1, world):
            other = union[r * n:(r + 1) * n]
            lo, hi = other.min(0).values - h, other.max(0).values + h
            near_mask |= ((cloud >= lo) & (cloud <= hi)).all(1)
        near = near_mask.nonzero()[:, 0]
        if near.numel() > 2_000_000:
            near = near[:2_000_000]
        ref = multiscale.process_single_core(cloud[near].contiguous(), union, edges, radii, out_dtype=np.float32)
        identical = bool(torch.equal(ref, out[near]))
        result = {"rows": int(near.numel()), "identical": identical,
                  "how": "rank 0's queries inside another tile's box grown by h = %.2f, recomputed on one GPU against "
                         "the union of all %d tiles" % (h, world)}
    del union
    return result


C4_EDGES = (0.5, 1.0, 2.0, 4.0, 8.0)
C4_RADII = (1.5, 3.0, 6.0, 12.0, 24.0)


def run_config4(args, nd, multiscale, dist, synth, lib, rank, world, dev):
    """BASELINE configs[3]: a 100M-point aerial tile, 5 scales (edge 0.5..8 m, r = 3e), split into `world` spatial
    tiles with halo exchange: STRONG scaling (the total is fixed).  N = 1 runs the whole tile on one GPU."""
    import ctypes
    import torch
    from nimrud_b200 import _lib
    total = args.config4_points
    n = total // world
    extent = math.sqrt(n / 8.0)
    cols = 2 if world >= 2 else 1
    cloud = synth.aerial_tile(n, seed=22 + rank, device=dev, origin=((rank % cols) * extent, (rank // cols) * extent))
    out = torch.empty((n, 4 * len(C4_RADII)), dtype=torch.float32, device=dev)
    edges_arr, edges_p = _lib.f64_array(C4_EDGES)
    radii_arr, radii_p = _lib.f64_array(C4_RADII)

    def step():
        if world > 1:
            nd.process_tile(cloud, C4_EDGES, C4_RADII, out=out, gather=False)
        else:
            _lib.check(lib.nbr_multiscale_features(
                ctypes.c_void_p(cloud.data_ptr()), _lib.F32, n, ctypes.c_void_p(cloud.data_ptr()), _lib.F32, n,
                edges_p, radii_p, len(C4_RADII), ctypes.c_void_p(out.data_ptr()), _lib.F32, 0, None, None,
                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    steps = max(1, min(args.steps, 5))
    for _ in range(3):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms /= steps
    block = {"workload": "BASELINE configs[3]: %dM-point synthetic aerial-LiDAR tile, 5 scales (edge 0.5..8 m, r = 3e), "
                         "%d spatial tile(s) with halo exchange" % (total // 1_000_000, world),
             "scaling": "strong", "points_total": n * world, "n_gpus": world, "ms_per_step": ms, "steps": steps,
             "value": n * world * len(C4_RADII) / (ms * 1e-3), "unit": UNIT}
    if world > 1:
        block["multi_gpu_check"] = seam_check(nd, multiscale, dist, cloud, out, C4_EDGES, C4_RADII, rank, world)
    del cloud, out
    torch.cuda.empty_cache()
    return block


def _timed_ms(fn, reps, torch):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_extras(args, lib, multiscale, synth, cloud, dev):
    """the other BASELINE configs that fit one GPU, device-resident, CUDA events: configs[0] (the reference's own
    example: r/e = 5, the 11-wide-window kernel), the equal-edge variant of configs[1] (one lattice, five
    radii: 7-wide, 11-wide and interval kernels), configs[2] (10M kNN, k = 10/20/50, ties) and the radix sort."""
    import ctypes
    import torch
    from nimrud_b200 import _lib
    from nimrud_b200.geometry import VoxelFilter
    extras = {}
    n = int(cloud.shape[0])
    # ---- configs[0]
    c0 = torch.from_numpy(synth.uniform_box()).to(dev)
    e0, r0 = (0.1, 0.2, 0.4), (0.5, 1.0, 2.0)
    ms = _timed_ms(lambda: multiscale.process_single_core(c0, c0, e0, r0, out_dtype=np.float32), 20, torch)
    extras["config0"] = {"workload": "BASELINE configs[0]: 100k uniform points in 20 x 20 x 2, edges (0.1, 0.2, 0.4), radii (0.5, 1, 2) "
                                     "(nimrud/minimal/multiscale.py example, r/e = 5)",
                         "ms_per_step": ms, "value": c0.shape[0] * 3 / (ms * 1e-3), "unit": UNIT,
                         "reference_1core": "5 723 points*scales/s measured on the reference itself (SURVEY.md 6)"}
    # ---- equal-edge variant of configs[1]
    ee, er = (0.2,) * 5, (0.4, 0.6, 0.8, 1.0, 1.2)
    out = torch.empty((n, 20), dtype=torch.float32, device=dev)
    ms = _timed_ms(lambda: multiscale.process_single_core(cloud, cloud, ee, er, out_dtype=np.float32, out=out), max(2, min(args.steps, 5)), torch)
    extras["equal_edge"] = {"workload": "configs[1] scene, one lattice (edge 0.2) shared by five radii 0.4 .. 1.2 (r/e = 2 .. 6)",
                            "ms_per_step": ms, "value": n * 5 / (ms * 1e-3), "unit": UNIT,
                            "mean_population": [round(float(out[:, 4 * s].mean()), 1) for s in range(5)]}
    del out
    # ---- radix sort of the voxel addresses of the kNN index (the sort inside np.unique, utils/geometry.py:150)
    kcloud = synth.urban_scene(n, seed=21, device=dev)
    vf = VoxelFilter(kcloud, 0.1)
    bits = int(sum(vf.widths))
    keys0 = vf.coordinate_to_address(kcloud).to(torch.int64).contiguous()
    keys, tmp = keys0.clone(), torch.empty_like(keys0)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def sort_once():
        keys.copy_(keys0)
        _lib.check(lib.nbr_sort_u64(ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(tmp.data_ptr()), n, 0, bits, stream))
    ms_copy = _timed_ms(lambda: keys.copy_(keys0), 5, torch)
    ms = _timed_ms(sort_once, 5, torch) - ms_copy
    passes = (bits + 7) // 8
    peak, _ = hbm_peak()
    extras["sort"] = {"what": "LSD radix sort of %d 64-bit voxel addresses, %d key bits, %d passes of 8 bits" % (n, bits, passes),
                      "ms": ms, "keys_per_s": n / (ms * 1e-3), "algorithmic_gbs": n * 16 * passes / ms / 1e6,
                      "frac_of_hbm_peak": n * 16 * passes / ms / 1e6 / peak,
                      "bytes": "16 B per key and pass (one read + one write); the histogram pass re-reads the keys"}
    del keys, tmp, keys0
    # ---- configs[2]: kNN over the voxel centres at e = 0.1, 1 % duplicated + 1 % lattice-aligned queries (ties)
    q = synth.with_ties(kcloud, 0.1, seed=21, fraction=0.01)
    index = None
    ms_build = _timed_ms(lambda: multiscale.LatticeIndex(kcloud, 0.1, indexed=True).close(), 2, torch)
    index = multiscale.LatticeIndex(kcloud, 0.1, indexed=True)
    res = {}
    ms50 = _timed_ms(lambda: res.__setitem__("r", index.knn(q, 50, ks=(10, 20, 50), out_dtype=np.float32)), 2, torch)
    idx, d2, _ = res["r"]
    # check a spatial subset against a float64 brute force on the device (total order (d2, index)); the CPU oracle check
    # of the same path is tests/test_knn_gpu.py and scripts/config3.py
    addr, cen = index.addresses_and_centres()
    lo = torch.tensor([120.0, 120.0], device=dev, dtype=torch.float64)
    sub = ((q[:, :2].double() >= lo) & (q[:, :2].double() < lo + 34.0)).all(1).nonzero()[:, 0][:20000]
    near = ((cen[:, :2] >= lo - 6.0) & (cen[:, :2] < lo + 40.0)).all(1).nonzero()[:, 0]
    ok_rows, checked = 0, 0
    if sub.numel() and near.numel() >= 50:
        dmax = d2[sub][:, -1].max().item() ** 0.5
        qs = q[sub].double()
        cl = cen[near]
        same_all = True
        for a in range(0, qs.shape[0], 256):
            blk = qs[a:a + 256]
            dd = ((blk[:, None, 0] - cl[None, :, 0]) ** 2 + (blk[:, None, 1] - cl[None, :, 1]) ** 2) + (blk[:, None, 2] - cl[None, :, 2]) ** 2
            # total order (d2, index): stable sort by index first (already ascending), then by d2
            order = torch.sort(dd, dim=1, stable=True).indices[:, :50]
            ref_idx = near[order]
            got = idx[sub[a:a + 256]].long()
            good = (got == ref_idx).all(1) & (d2[sub[a:a + 256]] == torch.gather(dd, 1, order)).all(1)
            ok_rows += int(good.sum())
            checked += int(blk.shape[0])
        check = {"queries": checked, "identical": bool(ok_rows == checked), "margin_ok": bool(dmax < 6.0),
                 "how": "float64 brute force on the device over the voxel centres around a 34 m x 34 m window, "
                        "stable sort = (d2, index) order"}
    else:
        check = None
    extras["knn"] = {"workload": "BASELINE configs[2]: %d queries (10M points + 1 %% duplicates + 1 %% lattice-aligned) against the %d voxel "
                                 "centres of the cloud at e = 0.1, k = 50 with features for k = 10 / 20 / 50" % (q.shape[0], index.n_voxels),
                     "ms": ms50, "queries_per_s": q.shape[0] / (ms50 * 1e-3), "value": q.shape[0] * 3 / (ms50 * 1e-3), "unit": UNIT,
                     "indexed_lattice_build_ms": ms_build, "check": check}
    index.close()
    del idx, d2, res
    # ---- configs[2], raw-point variant (the legacy sspedge = 0): the same queries against the 10M unfiltered points
    res = {}
    ms_raw = _timed_ms(lambda: res.__setitem__("r", multiscale.knn_points(q, kcloud, 50, ks=(10, 20, 50), out_dtype=np.float32)), 2, torch)
    idx, d2, _ = res["r"]
    near = ((kcloud[:, :2].double() >= lo - 6.0) & (kcloud[:, :2].double() < lo + 40.0)).all(1).nonzero()[:, 0]
    check = None
    if sub.numel() and near.numel() >= 50:
        dmax = d2[sub][:, -1].max().item() ** 0.5
        qs, cl = q[sub].double(), kcloud[near].double()
        ok_rows = checked = 0
        for a in range(0, qs.shape[0], 128):
            blk = qs[a:a + 128]
            dd = ((blk[:, None, 0] - cl[None, :, 0]) ** 2 + (blk[:, None, 1] - cl[None, :, 1]) ** 2) + (blk[:, None, 2] - cl[None, :, 2]) ** 2
            order = torch.sort(dd, dim=1, stable=True).indices[:, :50]
            good = (idx[sub[a:a + 128]].long() == near[order]).all(1) & (d2[sub[a:a + 128]] == torch.gather(dd, 1, order)).all(1)
            ok_rows += int(good.sum())
            checked += int(blk.shape[0])
        check = {"queries": checked, "identical": bool(ok_rows == checked), "margin_ok": bool(dmax < 6.0),
                 "how": "float64 brute force on the device over the raw points around a 34 m x 34 m window, stable sort = (d2, index) order"}
    extras["knn_raw_points"] = {"workload": "the same %d queries against the %d UNFILTERED points (search over raw points, legacy sspedge = 0), "
                                            "k = 50 with features for k = 10 / 20 / 50; cell index built inside the call" % (q.shape[0], n),
                                "ms": ms_raw, "queries_per_s": q.shape[0] / (ms_raw * 1e-3), "value": q.shape[0] * 3 / (ms_raw * 1e-3),
                                "unit": UNIT, "check": check}
    return extras


def gpu_arm(args, rank, world, local_rank):
    import ctypes
    import torch
    from nimrud_b200 import _lib, multiscale, synth

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    lib = _lib.lib()
    n = args.points
    ns = len(RADII)
    if world > 1:
        from nimrud_b200 import distributed as nd
        extent = math.sqrt(n / 40.0)
        cols = 2 if world >= 2 else 1
        origin = ((rank % cols) * extent, (rank // cols) * extent)
        cloud = synth.urban_scene(n, seed=20 + rank, device=dev, origin=origin)
    else:
        cloud = synth.urban_scene(n, seed=20, device=dev)
    out = torch.empty((n, 4 * ns), dtype=torch.float32, device=dev)
    counts = np.zeros(ns, dtype=np.int64)

    edges_arr, edges_p = _lib.f64_array(EDGES)
    radii_arr, radii_p = _lib.f64_array(RADII)

    def step(want_counts=False):
        if world > 1:
            return nd.process_tile(cloud, EDGES, RADII, out=out, gather=False, voxel_counts=counts if want_counts else None)
        cp = counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if want_counts else None
        _lib.check(lib.nbr_multiscale_features(
            ctypes.c_void_p(cloud.data_ptr()), _lib.F32, n, ctypes.c_void_p(cloud.data_ptr()), _lib.F32, n,
            edges_p, radii_p, ns, ctypes.c_void_p(out.data_ptr()), _lib.F32, 0, None, cp,
            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    step(want_counts=True)                               # voxel counts for the roofline's rho_s
    torch.cuda.synchronize()

    # ---- timed region: device-resident inputs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lib.nbr_timing_enable(1)
    phases = (ctypes.c_double * 8)()
    lib.nbr_timing_read(phases)                          # clear
    launches0 = lib.nbr_kernel_launches()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = lib.nbr_kernel_launches() - launches0
    lib.nbr_timing_read(phases)
    lib.nbr_timing_enable(0)
    if dist is not None:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * n * ns / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, host rows out, copies inside the timed region.  headline = the reference-facing host
    # entry point (C ABI) with pinned buffers and float64 rows (the reference's output dtype); beside it the same call
    # with float32 rows and the documented drop-in `process_single_core(numpy, numpy)` with plain (pageable) arrays
    e2e = None
    if not args.no_e2e:
        host_in = cloud.cpu().pin_memory()
        e2e_steps = max(1, min(args.steps, 3))
        results = {}
        legs = [("f64", torch.float64, _lib.F64), ("f32", torch.float32, _lib.F32)]
        for label, tdt, code in legs:
            host_out = torch.empty((n, 4 * ns), dtype=tdt).pin_memory()
            if world > 1:
                np_dt = np.float64 if tdt == torch.float64 else np.float32

                def host_step():
                    # host tile -> device, box table + halo push + lattices, rows -> host in batches (nbr_tile_step_host)
                    nd.process_tile_host(host_in, EDGES, RADII, out=host_out, out_dtype=np_dt, device=dev)
            else:
                def host_step():
                    _lib.check(lib.nbr_multiscale_features_host(
                        ctypes.c_void_p(host_in.data_ptr()), _lib.F32, n, ctypes.c_void_p(host_in.data_ptr()), _lib.F32,
                        n, edges_p, radii_p, ns, ctypes.c_void_p(host_out.data_ptr()), code, 0, None))
            host_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                host_step()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            results[label] = (world * n * ns * e2e_steps / dt, host_out.numel() * host_out.element_size())
            if label == "f64":
                check = host_out[:4096].to(torch.float32)
                assert torch.equal(check[:, 0::4], out[:4096, 0::4].cpu()), "e2e and device-resident results differ"
            del host_out
        e2e = {"value": results["f64"][0], "unit": UNIT, "h2d_bytes_per_step": int(host_in.numel() * 4),
               "d2h_bytes_per_step": int(results["f32"][1]), "out_dtype": "float64 (drop-in default)",
               "wire": "float32 rows over PCIe, widened to float64 by host threads inside the call; batches whose turn comes while "
                       "the host threads are behind are widened on the device and land as float64 (pinned result)",
               "value_float32_out": results["f32"][0], "d2h_bytes_per_step_float32_out": int(results["f32"][1]),
               "steps": e2e_steps, "timer": "host wall clock around the synchronous host-buffer call"}
        if world == 1:
            # the documented drop-in: numpy in, fresh numpy float64 out (pageable on both sides)
            np_in = host_in.numpy().copy()
            # warm-up in the loop's own pattern (the previous result is alive while the next call runs, so the loop
            # cycles through two recycled result buffers: both have to exist before the clock starts)
            for _ in range(3):
                res = multiscale.process_single_core(np_in, np_in, EDGES, RADII)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                res = multiscale.process_single_core(np_in, np_in, EDGES, RADII)
            dt = time.perf_counter() - t0
            assert res.dtype == np.float64 and np.array_equal(res[:4096, 0::4].astype(np.float32), out[:4096, 0::4].cpu().numpy())
            e2e["value_numpy_shim"] = n * ns * e2e_steps / dt
            e2e["numpy_shim"] = "nimrud_b200.multiscale.process_single_core(ndarray, ndarray): pageable input staged through " \
                                "pinned rings, fresh pageable float64 result"
            del res, np_in

    # ---- N > 1: the step with the final feature all-gather, and the seam check
    with_gather, seam = None, None
    if world > 1:
        g_steps = max(1, min(args.steps, 5))
        gathered = nd.process_tile(cloud, EDGES, RADII, out=out, gather="nccl")
        barrier()
        ev0.record()
        for _ in range(g_steps):
            gathered = nd.process_tile(cloud, EDGES, RADII, out=out, gather="nccl")
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        g_ms = float(t.item()) / g_steps
        with_gather = {"ms_per_step": g_ms, "value": world * n * ns / (g_ms * 1e-3), "unit": UNIT, "steps": g_steps,
                       "gathered_bytes_per_rank": int(gathered.numel() * gathered.element_size()),
                       "how": "all_gather_into_tensor of the (n, 20) float32 rows into one preallocated (N*n, 20) result on every rank"}
        same = bool(torch.equal(gathered[rank * n:(rank + 1) * n], out))
        # the same result without a collective call: the fused feature kernel stores every finished row into a staging
        # buffer of every other rank over NVLink, the receivers put the rows in place (process_tile(gather=True) ->
        # nbr_tile_step_gather)
        peer_out = torch.empty_like(gathered)
        peer = nd.process_tile(cloud, EDGES, RADII, gather="peer", out_all=peer_out)
        peer_same = bool(torch.equal(peer, gathered))
        del gathered
        barrier()
        ev0.record()
        for _ in range(g_steps):
            peer = nd.process_tile(cloud, EDGES, RADII, gather="peer", out_all=peer_out)
        ev1.record()
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p_ms = float(t.item()) / g_steps
        with_gather["nccl_ms_per_step"] = g_ms
        with_gather["nccl_value"] = with_gather["value"]
        with_gather["peer_ms_per_step"] = p_ms
        with_gather["peer_value"] = world * n * ns / (p_ms * 1e-3)
        # ms_per_step / value: what process_tile(gather=True) does by default at this world size
        with_gather["default"] = "peer" if world <= 4 else "nccl"
        if world <= 4:
            with_gather["ms_per_step"] = p_ms
            with_gather["value"] = with_gather["peer_value"]
        with_gather["identical_to_nccl_gather"] = peer_same
        with_gather["how"] = "peer: finished rows stored into a staging buffer of every other rank by the feature kernel itself " \
                             "(peer-mapped memory over NVLink, a warp's 32 rows as one contiguous piece + row numbers), " \
                             "stream-ordered signal + wait, receivers put the rows in place in the preallocated (N*n, 20) " \
                             "float32 result; no collective call.  nccl: the same step followed by all_gather_into_tensor"
        del peer, peer_out
        seam = seam_check(nd, multiscale, dist, cloud, out, EDGES, RADII, rank, world)
        if seam is not None:
            seam["gathered_rows_match_local"] = same

    extras = None
    if world == 1 and not args.no_extras:
        extras = run_extras(args, lib, multiscale, synth, cloud, dev)

    # ---- BASELINE configs[3]: 100M-point aerial tile split over the ranks (strong scaling)
    config4 = None
    if not args.no_config4:
        config4 = run_config4(args, nd if world > 1 else None, multiscale, dist, synth, lib, rank, world, dev)

    # the sampler covers the device-resident timed region and the e2e region (both under load)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        vox = torch.from_numpy(counts.astype(np.int64)).to(dev)
        dist.all_reduce(vox)
        counts_all = vox.cpu().numpy()
    else:
        counts_all = counts
    if rank != 0:
        if dist is not None:
            nd.release_mailboxes()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (fused feature kernels, all scales of one step)
    peak, peak_src = hbm_peak()
    feat_ms = phases[3] / args.steps
    # rho_s = voxels a rank's lattices hold (tile + halo at N > 1, summed over the ranks) / queries; bytes per rank
    rho = counts_all / float(n * world)
    algo_bytes = float(sum(n * (28.0 + 12.0 * r) for r in rho))
    achieved = algo_bytes / (feat_ms * 1e-3) / 1e9 if feat_ms > 0 else 0.0
    traffic, traffic_src, secondary = None, None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_rows3_traffic.json")))
        if n == 10_000_000:
            traffic = float(tr["dram_bytes_read"] + tr["dram_bytes_write"])
            traffic_src = "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel " \
                          "on this workload (profiles/r02_rows3_traffic.json)"
            # the secondary bounds beside the HBM figure.  the kernel is bound by instruction issue
            secondary = {"bound": "instruction issue (integer / FP32 pipes)",
                         "issue_slot_util": tr["issue_slot_util_pct"] / 100.0,
                         "inst_per_point_scale": tr["warp_instructions"] / float(n * ns),
                         "inst_unit": "warp instructions per query and scale (x 32 / threads_per_instruction for thread instructions)",
                         "threads_per_instruction": tr["threads_per_instruction"],
                         "l2_gbs": tr["lts_sectors"] * 32.0 / (tr["gpu_time_ms_under_ncu"] * 1e-3) / 1e9,
                         "l2_hit_rate": tr["l2_hit_rate_pct"] / 100.0,
                         "pipes_pct_of_peak": tr["pipes_pct_of_peak"],
                         "source": "the same ncu capture (profiles/r02_rows3_ncu_raw.csv)"}
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "nbr::rows3_kernel: fused radius query + covariance + eigen + features "
                                          "(one launch for all scales of a step)",
                "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "secondary": secondary,
                "algorithmic_bytes_per_step": algo_bytes, "kernel_ms_per_step": feat_ms,
                "bytes_per_point_scale": "28 + 12*rho_s (SURVEY.md 8d); rho_s = unique voxels / queries = %s"
                                         % [round(float(r), 4) for r in rho],
                "phase_ms_per_step": {"bbox": phases[0] / args.steps, "index": phases[1] / args.steps,
                                      "order": phases[2] / args.steps, "features": feat_ms,
                                      "tile_boxes": phases[4] / args.steps, "halo_push": phases[5] / args.steps,
                                      "halo_wait_inside_index": phases[6] / args.steps}}

    cpu_baseline = None
    if world == 1 and not args.no_cpu:
        workers = max(1, min(os.cpu_count() or 1, 64))
        v, ms, sample = run_cpu_reference(1, 0, workers)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample,
                        "ms": ms}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 moments + f64 membership/eigen, f32 out", "data": "synthetic",
        "config": workload_config(n, world), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "voxels_per_scale": [int(c) for c in counts_all],
    }
    if world > 1:
        line["with_gather"] = with_gather
        line["multi_gpu_check"] = seam
        line["halo_transport"] = os.environ.get("NBR_HALO", "mailbox") + \
            (": peer-mapped mailboxes over NVLink (csrc/mailbox.cu), one host synchronisation per step"
             if os.environ.get("NBR_HALO", "mailbox") != "nccl" else ": NCCL all-gather + all-to-all-v")
    if config4 is not None:
        line["config4"] = config4
    if extras is not None:
        line["extra"] = extras
    print(json.dumps(line), flush=True)
    if dist is not None:
        nd.release_mailboxes()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-config4", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--config4-points", type=int, default=100_000_000)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
