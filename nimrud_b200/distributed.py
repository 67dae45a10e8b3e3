"""
multi-GPU partitioning of the multiscale eigenfeature path: one process per GPU, each owning a
spatial tile of the cloud (query == search inside the tile).

    1. all-reduce (min/max) of the tile bounding boxes -> the global box.  every rank anchors its voxel
       lattices on it, so a voxel is the same voxel on every GPU (the precedent in the
       reference is nested_regions, nimrud/utils/geometry.py:203-253, and Partitions with
       buffer = largest scale, nimrud/prototypes/mso.py:286,317).
    2. halo exchange (all-to-all-v over NCCL): rank r receives every foreign point within
       h = max_s(r_s + e_s / 2) (per axis) of its tile box.  a voxel centre within r_s of a query of the
       tile holds a point within r_s + e_s/2 per axis of that query, so tile + halo reproduces exactly
       the voxels the unpartitioned run would see around every query of the tile.
    3. the single-GPU path on (queries = tile, search = tile + halo, lattice anchored globally).
    4. optional all-gather-v of the feature rows.

There is no data-path collective besides 2 and 4.  The host logic (box reduction, halo selection,
exchange) is backend-agnostic torch.distributed, so it is tested on CPU with gloo; the compute
function is the CUDA path unless a test injects another one.
"""
import numpy as np
import torch
import torch.distributed as dist


def halo_width(edge_lengths, radii):
    """per-axis halo that makes a tile self-sufficient: max over scales of r + e/2 (plus float slack)."""
    return max(float(r) + float(e) / 2 for e, r in zip(edge_lengths, radii)) * (1 + 1e-6)


def tile_box(cloud):
    """(lo, hi) float64 tensors (3,) on the cloud's device (min / max are exact in the cloud's own dtype)."""
    if cloud.is_cuda and cloud.dtype in (torch.float32, torch.float64) and cloud.is_contiguous() and cloud.shape[0] > 0:
        from . import _lib
        from ._util import ptr, stream_ptr
        box = torch.empty(6, dtype=torch.float64, device=cloud.device)
        with torch.cuda.device(cloud.device):
            _lib.check(_lib.lib().nbr_bbox(ptr(cloud), _lib.F32 if cloud.dtype == torch.float32 else _lib.F64,
                                           int(cloud.shape[0]), 3, ptr(box), stream_ptr(cloud.device)))
        return box[:3], box[3:]
    return cloud.min(0).values.to(torch.float64), cloud.max(0).values.to(torch.float64)


def global_box(lo, hi, group=None):
    lo = lo.clone(); hi = hi.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return lo, hi


def select_halo(cloud, box_lo, box_hi, h):
    """indices of the points of `cloud` inside [box_lo - h, box_hi + h] (inclusive, all three axes).
    the bounds are rounded OUTWARD to the cloud's dtype, so the test runs on the coordinates as stored
    (no float64 copy of the cloud) and can only select a superset of the exact float64 test."""
    lo = (box_lo - h).to(torch.float64)
    hi = (box_hi + h).to(torch.float64)
    if cloud.dtype != torch.float64:
        lo_c, hi_c = lo.to(cloud.dtype), hi.to(cloud.dtype)
        lo_c = torch.where(lo_c.to(torch.float64) > lo, torch.nextafter(lo_c, torch.full_like(lo_c, -float("inf"))), lo_c)
        hi_c = torch.where(hi_c.to(torch.float64) < hi, torch.nextafter(hi_c, torch.full_like(hi_c, float("inf"))), hi_c)
        lo, hi = lo_c, hi_c
    inside = ((cloud >= lo) & (cloud <= hi)).all(1)
    return inside.nonzero(as_tuple=True)[0]


def _count_halos_cuda(cloud, grown_boxes):
    """first pass of the CUDA halo selection: per destination box (lo, hi float64 lists, already grown) the number
    of points of `cloud` inside it, as a DEVICE tensor (no host synchronisation), plus what the fill pass needs."""
    import ctypes
    from . import _lib
    from ._util import ptr, stream_ptr
    lib = _lib.lib()
    code = _lib.F32 if cloud.dtype == torch.float32 else _lib.F64
    n = int(cloud.shape[0])
    state = []
    counts = []
    with torch.cuda.device(cloud.device):
        s = stream_ptr(cloud.device)
        for first in range(0, len(grown_boxes), 8):
            chunk = grown_boxes[first:first + 8]
            flat = np.ascontiguousarray([list(lo) + list(hi) for lo, hi in chunk], dtype=np.float64).reshape(-1)
            boxes_p = flat.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
            cnt = torch.zeros(2 * len(chunk), dtype=torch.int64, device=cloud.device)     # counts | cursors
            _lib.check(lib.nbr_halo_count(ptr(cloud), code, n, boxes_p, len(chunk), ptr(cnt), s))
            state.append((flat, cnt, len(chunk)))
            counts.append(cnt[:len(chunk)])
    return (torch.cat(counts) if len(counts) != 1 else counts[0]), state


def _fill_halos_cuda(cloud, state, counts_host):
    """second pass: the selected points of every destination, destination by destination -> (m, 3)."""
    import ctypes
    from . import _lib
    from ._util import ptr, stream_ptr
    lib = _lib.lib()
    code = _lib.F32 if cloud.dtype == torch.float32 else _lib.F64
    n = int(cloud.shape[0])
    parts = []
    at = 0
    with torch.cuda.device(cloud.device):
        s = stream_ptr(cloud.device)
        for flat, cnt, m in state:
            c = counts_host[at:at + m]
            at += m
            offs = np.concatenate([[0], np.cumsum(c)[:-1]]).astype(np.int64)
            buf = torch.empty((int(sum(c)), 3), dtype=cloud.dtype, device=cloud.device)
            if sum(c):
                _lib.check(lib.nbr_halo_fill(ptr(cloud), code, n, flat.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), m,
                                             offs.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                             ptr(cnt[m:]), ptr(buf), s))
            parts.append(buf)
    return torch.cat(parts, 0) if len(parts) != 1 else parts[0]


def gather_boxes(cloud, group=None):
    """(world, 6) float64 CPU tensor of every rank's tile box [lo, hi], and this rank's (lo, hi) device tensors."""
    world = dist.get_world_size(group)
    lo, hi = tile_box(cloud)
    boxes = [torch.empty(6, dtype=torch.float64, device=cloud.device) for _ in range(world)]
    dist.all_gather(boxes, torch.cat([lo, hi]), group=group)
    return torch.stack(boxes).cpu(), (lo, hi)                  # one small device->host copy


def exchange_halo(cloud, edge_lengths, radii, group=None, gathered=None):
    """
    -> (halo points received from the other ranks (m,3), same dtype/device as cloud,
        global (lo, hi) float64 CPU tensors, own tile (lo, hi))
    gathered: the result of gather_boxes(cloud, group) if the caller already has it.
    """
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    all_boxes, (lo, hi) = gathered if gathered is not None else gather_boxes(cloud, group)
    g_lo = all_boxes[:, :3].min(0).values                      # == all-reduce(min/max) of the tile boxes (host copy)
    g_hi = all_boxes[:, 3:].max(0).values
    h = halo_width(edge_lengths, radii)
    my_lo, my_hi = all_boxes[rank, :3], all_boxes[rank, 3:]

    # tiles whose grown box misses this tile's box get nothing
    targets = [dst for dst in range(world)
               if dst != rank and not bool(((all_boxes[dst, :3] - h) > my_hi).any())
               and not bool(((all_boxes[dst, 3:] + h) < my_lo).any())]
    send_counts = [0] * world
    if cloud.is_cuda and cloud.dtype in (torch.float32, torch.float64) and cloud.is_contiguous():
        # count on the device, exchange the counts device to device, and read both directions with ONE host
        # synchronisation; the fill pass and the point exchange follow
        counts_t = torch.zeros(world, dtype=torch.int64, device=cloud.device)
        state = None
        if targets:
            grown = [((all_boxes[d, :3] - h).tolist(), (all_boxes[d, 3:] + h).tolist()) for d in targets]
            cnt_dev, state = _count_halos_cuda(cloud, grown)
            counts_t[torch.tensor(targets, device=cloud.device)] = cnt_dev
        recv_counts_t = torch.empty_like(counts_t)
        dist.all_to_all_single(recv_counts_t, counts_t, group=group)
        both = torch.stack([counts_t, recv_counts_t]).tolist()
        send_counts, recv_counts = [int(v) for v in both[0]], [int(v) for v in both[1]]
        send_buf = _fill_halos_cuda(cloud, state, [send_counts[d] for d in targets]) if targets else cloud[:0]
    else:
        # host logic on CPU tensors (gloo tests): same inclusive selection with torch ops
        send_parts = []
        for dst in targets:
            idx = select_halo(cloud, all_boxes[dst, :3].to(cloud.device), all_boxes[dst, 3:].to(cloud.device), h)
            send_parts.append(cloud[idx])
            send_counts[dst] = int(idx.numel())
        send_buf = torch.cat(send_parts, 0) if send_parts else cloud[:0]
        counts_t = torch.tensor(send_counts, dtype=torch.int64, device=cloud.device)
        recv_counts_t = torch.empty_like(counts_t)
        dist.all_to_all_single(recv_counts_t, counts_t, group=group)
        recv_counts = [int(v) for v in recv_counts_t.tolist()]
    send_buf = send_buf.contiguous().reshape(-1)
    recv_buf = torch.empty(sum(recv_counts) * 3, dtype=cloud.dtype, device=cloud.device)
    dist.all_to_all_single(recv_buf, send_buf, output_split_sizes=[3 * c for c in recv_counts],
                           input_split_sizes=[3 * c for c in send_counts], group=group)
    return recv_buf.reshape(-1, 3), (g_lo, g_hi), (lo, hi)


def _gpu_compute(query, search, edge_lengths, radii, bbox, out_dtype, out):
    from . import multiscale
    return multiscale.process_single_core(query, search, edge_lengths, radii, out_dtype=out_dtype,
                                          global_bbox=bbox, out=out)


_side_streams = {}


def _side_stream(device):
    # one persistent side stream per device: the caching allocator keeps per-stream pools, a fresh stream per
    # call would allocate its buffers anew every time
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def _process_tile_cuda(cloud, edge_lengths, radii, out, out_dtype, group):
    """the CUDA tile path: the tile is ordered on a side stream while the halo exchange is in flight; the
    lattices are then built from the ordered tile + the received halo points, the tile's points are the queries."""
    import ctypes
    from . import _lib
    from ._util import ptr, stream_ptr
    from .multiscale import _out_code, _TORCH_OUT
    lib = _lib.lib()
    code = _lib.F32 if cloud.dtype == torch.float32 else _lib.F64
    n = int(cloud.shape[0])
    rank = dist.get_rank(group)
    gathered = gather_boxes(cloud, group)
    all_boxes = gathered[0]
    h = halo_width(edge_lengths, radii)
    g_lo, g_hi = all_boxes[:, :3].min(0).values, all_boxes[:, 3:].max(0).values
    # every halo point this rank can receive lies inside its own box grown by h
    local_lo = torch.maximum(all_boxes[rank, :3] - h, g_lo)
    local_hi = torch.minimum(all_boxes[rank, 3:] + h, g_hi)
    f64p = ctypes.POINTER(ctypes.c_double)
    glob = np.ascontiguousarray(torch.cat([g_lo, g_hi]).numpy(), dtype=np.float64)
    local = np.ascontiguousarray(torch.cat([local_lo, local_hi]).numpy(), dtype=np.float64)
    mine = np.ascontiguousarray(all_boxes[rank].numpy(), dtype=np.float64)
    finest = float(min(edge_lengths))
    origin = np.zeros(3, dtype=np.float64)
    _lib.check(lib.nbr_brick_origin(glob.ctypes.data_as(f64p), local.ctypes.data_as(f64p), finest,
                                    origin.ctypes.data_as(f64p)))
    np_out, out_code = _out_code(out_dtype)
    n_scales = len(radii)
    edges_arr, edges_p = _lib.f64_array(list(edge_lengths))
    radii_arr, radii_p = _lib.f64_array(list(radii))
    with torch.cuda.device(cloud.device):
        main = torch.cuda.current_stream(cloud.device)
        side = _side_stream(cloud.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            perm = torch.empty(n, dtype=torch.int32, device=cloud.device)
            ordered = torch.empty_like(cloud)
            _lib.check(lib.nbr_order_cloud(ptr(cloud), code, n, mine.ctypes.data_as(f64p), origin.ctypes.data_as(f64p),
                                           finest, ptr(perm), ptr(ordered), ctypes.c_void_p(side.cuda_stream)))
        halo, _, _ = exchange_halo(cloud, edge_lengths, radii, group, gathered=gathered)
        main.wait_stream(side)
        perm.record_stream(main)
        ordered.record_stream(main)
        if out is None:
            out = torch.zeros((n, 4 * n_scales), dtype=_TORCH_OUT[np_out], device=cloud.device)
        _lib.check(lib.nbr_multiscale_features_tile(
            ptr(ordered), ptr(perm), code, n, ptr(halo) if halo.numel() else None, int(halo.shape[0]),
            local.ctypes.data_as(f64p), glob.ctypes.data_as(f64p), edges_p, radii_p, n_scales, ptr(out), out_code, 0,
            None, stream_ptr(cloud.device)))
    return out


def process_tile(cloud, edge_lengths, radii, out=None, out_dtype=np.float32, gather=False, group=None,
                 compute=None):
    """
    features of this rank's tile (n_local, 4*S).  `cloud`: (n_local, 3) tensor on this rank's device.
    gather=True: returns the rows of every rank, concatenated in rank order, on every rank.
    compute(query, search, edges, radii, (lo, hi) numpy, out_dtype, out) -> features; default = CUDA path.
    """
    assert len(edge_lengths) == len(radii), "edge_lengths and radii should be equal-length sequences."
    if (compute is None and cloud.is_cuda and cloud.dtype in (torch.float32, torch.float64) and cloud.is_contiguous()
            and cloud.shape[0] >= 2):
        feats = _process_tile_cuda(cloud, edge_lengths, radii, out, out_dtype, group)
    else:
        halo, (g_lo, g_hi), _ = exchange_halo(cloud, edge_lengths, radii, group)
        search = torch.cat([cloud, halo], 0) if halo.numel() else cloud
        bbox = (g_lo.cpu().numpy(), g_hi.cpu().numpy())
        # the queries are handed over as the first rows of the search buffer: the CUDA path then orders tile + halo
        # once, builds the lattices from the ordered copy and keeps only the tile's points as queries
        query = search[:cloud.shape[0]]
        feats = (compute or _gpu_compute)(query, search, edge_lengths, radii, bbox, out_dtype, out)
    if not gather:
        return feats
    world = dist.get_world_size(group)
    n_local = torch.tensor([feats.shape[0]], dtype=torch.int64, device=feats.device)
    sizes = [torch.empty_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    # all-gather-v: pad every contribution to the largest tile, gather, drop the padding
    cap = max(sizes)
    padded = feats.contiguous()
    if padded.shape[0] < cap:
        padded = torch.cat([padded, padded.new_zeros((cap - padded.shape[0], feats.shape[1]))], 0)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], 0)
