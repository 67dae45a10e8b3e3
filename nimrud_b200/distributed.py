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

There is no data-path collective besides 2 and 4.

Two transports for step 1 + 2:
  * CUDA tensors (the product path): HALO MAILBOXES over peer-mapped memory (csrc/mailbox.cu).  every rank's tile
    box goes into every peer's box table, the halo points are stored straight into the destination's
    mailbox by one pass over the tile (remote atomic cursor + remote stores over NVLink), flags signal
    completion.  no NCCL call, no count hand-shake, ONE host synchronisation per step (the boxes -- the
    single-GPU path has the same one for its bounding box).  the same code drives several tiles inside
    one process (`process_tiles_local`), which is how the single-GPU tests prove tile + halo == unpartitioned.
  * any other tensors (gloo tests on CPU, NBR_HALO=nccl): backend-agnostic torch.distributed collectives
    (all-gather of the boxes, all-to-all-v of counts and points).
"""
import ctypes
import os

import numpy as np
import torch
import torch.distributed as dist


def halo_width(edge_lengths, radii):
    """per-axis halo that makes a tile self-sufficient: max over scales of r + e/2 (plus float slack)."""
    return max(float(r) + float(e) / 2 for e, r in zip(edge_lengths, radii)) * (1 + 1e-6)


def tile_box(cloud):
    """(lo, hi) float64 tensors (3,) on the cloud's device (min / max are exact in the cloud's own dtype)."""
    if cloud.shape[0] == 0:
        inf = torch.full((3,), float("inf"), dtype=torch.float64, device=cloud.device)
        return inf, -inf                                           # the neutral box of min / max
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


# --------------------------------------------------------------------------------------------------
# halo mailboxes (csrc/mailbox.cu)
# --------------------------------------------------------------------------------------------------
def default_capacity(n_local):
    """rows a rank's mailbox can receive per step: NBR_HALO_CAPACITY, or max(4M, the tile's own size)."""
    env = os.environ.get("NBR_HALO_CAPACITY")
    return int(env) if env else max(4 << 20, int(n_local))


class HaloMailbox(object):
    """this rank's mailbox and its view of the peers' (one per (group, device, dtype); reused by every step)."""

    def __init__(self, rank, world, device, dtype, capacity_rows):
        from . import _lib
        self.rank, self.world, self.device, self.dtype = int(rank), int(world), torch.device(device), dtype
        self.capacity = int(capacity_rows)
        self.code = _lib.F32 if dtype == torch.float32 else _lib.F64
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_mailbox_create(ctypes.byref(handle), self.rank, self.world, self.code, self.capacity))
        self.handle = handle
        self.boxes = np.zeros((self.world, 8), dtype=np.float64)
        self.cache = {}

    def ipc_handle(self):
        from . import _lib
        buf = ctypes.create_string_buffer(64)
        _lib.check(_lib.lib().nbr_mailbox_ipc_handle(self.handle, buf))
        return bytes(buf.raw)

    @classmethod
    def connect_group(cls, device, dtype, capacity_rows, group=None):
        """collective: every rank creates its mailbox and maps every peer's through CUDA IPC."""
        from . import _lib
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        mb = cls(rank, world, device, dtype, capacity_rows)
        infos = [None] * world
        dist.all_gather_object(infos, (mb.ipc_handle(), mb.capacity), group=group)
        with torch.cuda.device(mb.device):
            for peer, (h, cap) in enumerate(infos):
                if peer == rank:
                    continue
                _lib.check(_lib.lib().nbr_mailbox_connect_ipc(mb.handle, peer, ctypes.c_char_p(h)))
                _lib.check(_lib.lib().nbr_mailbox_set_peer_capacity(mb.handle, peer, int(cap)))
        return mb

    @classmethod
    def local_set(cls, world, device, dtype, capacity_rows):
        """`world` mailboxes in this process (several tiles on one device, or one per visible device)."""
        from . import _lib
        devices = device if isinstance(device, (list, tuple)) else [device] * world
        boxes = [cls(r, world, devices[r], dtype, capacity_rows) for r in range(world)]
        for a in boxes:
            with torch.cuda.device(a.device):
                for b in boxes:
                    if a is not b:
                        _lib.check(_lib.lib().nbr_mailbox_connect_local(a.handle, b.rank, b.handle))
        return boxes

    # ---- staging buffers of the feature all-gather through peer stores
    def gather_bytes(self):
        from . import _lib
        nbytes = ctypes.c_uint64(0)
        _lib.lib().nbr_mailbox_gather_ptr(self.handle, ctypes.byref(nbytes))
        return int(nbytes.value)

    def gather_ipc_handle(self):
        from . import _lib
        buf = ctypes.create_string_buffer(64)
        _lib.check(_lib.lib().nbr_mailbox_gather_ipc_handle(self.handle, buf))
        return bytes(buf.raw)

    def ensure_gather(self, nbytes, group=None):
        """collective: every rank's staging buffer holds at least nbytes (the same value on every rank) and is mapped by
        every peer.  a no-op when the buffers are large enough already."""
        from . import _lib
        nbytes = int(nbytes)
        if self.gather_bytes() >= nbytes and nbytes > 0:
            return
        nbytes = max(nbytes, 256)
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_mailbox_disconnect(self.handle, 1))
        dist.barrier(group)                                  # nobody writes into or maps the buffers that are freed now
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_mailbox_gather_alloc(self.handle, nbytes))
            infos = [None] * self.world
            dist.all_gather_object(infos, self.gather_ipc_handle(), group=group)
            for peer, h in enumerate(infos):
                if peer != self.rank:
                    _lib.check(_lib.lib().nbr_mailbox_gather_connect_ipc(self.handle, peer, ctypes.c_char_p(h), nbytes))
        dist.barrier(group)

    @staticmethod
    def ensure_gather_local(mailboxes, nbytes):
        """the same for the mailboxes of one process (local_set)."""
        from . import _lib
        nbytes = max(int(nbytes), 256)
        if all(mb.gather_bytes() >= nbytes for mb in mailboxes):
            return
        for mb in mailboxes:
            torch.cuda.synchronize(mb.device)
        for mb in mailboxes:
            with torch.cuda.device(mb.device):
                _lib.check(_lib.lib().nbr_mailbox_gather_alloc(mb.handle, nbytes))
        for a in mailboxes:
            for b in mailboxes:
                if a is not b:
                    _lib.check(_lib.lib().nbr_mailbox_gather_connect_local(a.handle, b.rank, b.handle))

    def status(self):
        """{timeout, dropped, pushed}: valid after the device was synchronised."""
        from . import _lib
        st = (ctypes.c_uint64 * 3)()
        _lib.check(_lib.lib().nbr_mailbox_status(self.handle, st))
        return {"timeout": int(st[0]), "dropped": int(st[1]), "pushed": int(st[2])}

    def received(self):
        """(m, 3) tensor: the halo points of the last completed step (synchronises; for tests and debugging)."""
        from . import _lib
        from ._util import ptr, stream_ptr
        out = torch.empty((self.capacity, 3), dtype=self.dtype, device=self.device)
        n = ctypes.c_int64(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_mailbox_read(self.handle, ptr(out), self.capacity, ctypes.byref(n), stream_ptr(self.device)))
        return out[:n.value]

    def close(self):
        from . import _lib
        if getattr(self, "handle", None):
            with torch.cuda.device(self.device):
                torch.cuda.synchronize(self.device)
                _lib.lib().nbr_mailbox_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the four steps
    def publish(self, cloud):
        from . import _lib
        from ._util import ptr, stream_ptr
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_tile_box_publish(self.handle, ptr(cloud) if cloud.shape[0] else None, self.code,
                                                       int(cloud.shape[0]), stream_ptr(self.device)))

    def wait_boxes(self):
        """(world, 8) float64: lo, hi, n_points, 0 of every tile.  the step's one host synchronisation."""
        from . import _lib
        from ._util import stream_ptr
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_tile_boxes_wait(self.handle, self.boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                      stream_ptr(self.device)))
        return self.boxes

    def push(self, cloud, h):
        from . import _lib
        from ._util import ptr, stream_ptr
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_halo_push(self.handle, ptr(cloud) if cloud.shape[0] else None, self.code,
                                                int(cloud.shape[0]), self.boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                                float(h), stream_ptr(self.device)))


_mailboxes = {}


def _group_mailbox(cloud, group):
    key = (id(group) if group is not None else 0, cloud.device.index, cloud.dtype)
    mb = _mailboxes.get(key)
    if mb is None:
        mb = HaloMailbox.connect_group(cloud.device, cloud.dtype, default_capacity(cloud.shape[0]), group)
        _mailboxes[key] = mb
    return mb


def release_mailboxes():
    """collective: free the cached mailboxes (call before destroy_process_group).  every rank first closes its mappings
    of the peers' memory, the ranks meet, then the owners free."""
    from . import _lib
    boxes = list(_mailboxes.values())
    for mb in boxes:
        if getattr(mb, "handle", None):
            with torch.cuda.device(mb.device):
                _lib.check(_lib.lib().nbr_mailbox_disconnect(mb.handle, 0))
    if boxes and dist.is_available() and dist.is_initialized():
        dist.barrier()
    for mb in boxes:
        mb.close()
    _mailboxes.clear()


def _tile_geometry(boxes, rank, edge_lengths, radii):
    """host side of a step: global box, this tile's box grown by the halo width, the brick origin of the order."""
    from . import _lib
    h = halo_width(edge_lengths, radii)
    filled = boxes[:, 6] > 0
    if not filled.any():
        raise ValueError("need at least 2 points to define a voxel grid")
    g_lo = boxes[filled, 0:3].min(0)
    g_hi = boxes[filled, 3:6].max(0)
    glob = np.ascontiguousarray(np.concatenate([g_lo, g_hi]), dtype=np.float64)
    mine = np.ascontiguousarray(boxes[rank, :6], dtype=np.float64)
    local = np.ascontiguousarray(np.concatenate([np.maximum(mine[:3] - h, g_lo), np.minimum(mine[3:] + h, g_hi)]))
    origin = np.zeros(3, dtype=np.float64)
    if boxes[rank, 6] > 0:
        f64p = ctypes.POINTER(ctypes.c_double)
        _lib.check(_lib.lib().nbr_brick_origin(glob.ctypes.data_as(f64p), local.ctypes.data_as(f64p),
                                               float(min(edge_lengths)), origin.ctypes.data_as(f64p)))
    return h, glob, mine, local, origin


def _check_out(out, n, n_scales, np_out, device):
    from .multiscale import _TORCH_OUT
    if out is None:
        return torch.zeros((n, 4 * n_scales), dtype=_TORCH_OUT[np_out], device=device)
    if (tuple(out.shape) != (n, 4 * n_scales) or out.dtype != _TORCH_OUT[np_out] or not out.is_cuda
            or out.device != device or not out.is_contiguous()):
        raise ValueError("out has the wrong shape, dtype or device")
    return out


def _order_tile(cloud, mine, origin, finest, stream):
    """perm, ordered copy of the tile in the feature kernels' query order, on `stream`."""
    from . import _lib
    from ._util import ptr
    f64p = ctypes.POINTER(ctypes.c_double)
    n = int(cloud.shape[0])
    perm = torch.empty(n, dtype=torch.int32, device=cloud.device)
    ordered = torch.empty_like(cloud)
    code = _lib.F32 if cloud.dtype == torch.float32 else _lib.F64
    if n:
        _lib.check(_lib.lib().nbr_order_cloud(ptr(cloud), code, n, mine.ctypes.data_as(f64p), origin.ctypes.data_as(f64p),
                                              finest, ptr(perm), ptr(ordered), ctypes.c_void_p(stream.cuda_stream)))
    return perm, ordered


def _tile_features_mb(mb, ordered, perm, local, glob, edge_lengths, radii, out, out_code, voxel_counts=None):
    from . import _lib
    from ._util import ptr, stream_ptr
    f64p = ctypes.POINTER(ctypes.c_double)
    edges_arr, edges_p = _lib.f64_array(list(edge_lengths))
    radii_arr, radii_p = _lib.f64_array(list(radii))
    n = int(ordered.shape[0])
    _lib.check(_lib.lib().nbr_multiscale_features_tile_mb(
        ptr(ordered) if n else None, ptr(perm) if n else None, mb.code, n, mb.handle, local.ctypes.data_as(f64p),
        glob.ctypes.data_as(f64p), edges_p, radii_p, len(radii), ptr(out) if n else None, out_code, 0,
        voxel_counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if voxel_counts is not None and n else None,
        stream_ptr(ordered.device)))


def _process_tile_cuda(cloud, edge_lengths, radii, out, out_dtype, group, voxel_counts=None):
    """the CUDA tile path over halo mailboxes, one C call per step (nbr_tile_step): box table -> (host) -> push ->
    order -> lattices + features.  returns (features, boxes)."""
    from . import _lib
    from ._util import ptr, stream_ptr
    from .multiscale import _out_code
    np_out, out_code = _out_code(out_dtype)
    n = int(cloud.shape[0])
    out = _check_out(out, n, len(radii), np_out, cloud.device)
    mb = _group_mailbox(cloud, group)
    edges_arr, edges_p = _lib.f64_array(list(edge_lengths))
    radii_arr, radii_p = _lib.f64_array(list(radii))
    with torch.cuda.device(cloud.device):
        _lib.check(_lib.lib().nbr_tile_step(
            mb.handle, ptr(cloud) if n else None, mb.code, n, edges_p, radii_p, len(radii), ptr(out) if n else None,
            out_code, 0, mb.boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            voxel_counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if voxel_counts is not None and n else None,
            stream_ptr(cloud.device)))
    return out, mb.boxes


def _process_tile_gather(cloud, edge_lengths, radii, out_all, out_dtype, group, voxel_counts=None):
    """nbr_tile_step_gather: the tile step with the feature all-gather inside.  staging buffers and result are sized from
    the box table: the first step (and any step whose tiles outgrow them) reports NBR_ERR_CAPACITY on every rank alike,
    the buffers are re-made collectively and the step is repeated."""
    from . import _lib
    from ._util import ptr, stream_ptr
    from .multiscale import _out_code, _TORCH_OUT
    np_out, out_code = _out_code(out_dtype)
    n = int(cloud.shape[0])
    cols = 4 * len(radii)
    row_bytes = cols * np.dtype(np_out).itemsize
    mb = _group_mailbox(cloud, group)
    edges_arr, edges_p = _lib.f64_array(list(edge_lengths))
    radii_arr, radii_p = _lib.f64_array(list(radii))
    offsets = np.zeros(mb.world + 1, dtype=np.int64)
    given = out_all is not None
    if given and (out_all.dim() != 2 or out_all.shape[1] != cols or out_all.dtype != _TORCH_OUT[np_out] or not out_all.is_cuda
                  or not out_all.is_contiguous()):
        raise ValueError("out has the wrong shape, dtype or device")
    key = ("gather_total", cols, np.dtype(np_out).str)
    total = mb.cache.get(key, 0)
    for attempt in range(2):
        res = out_all if given else torch.empty((total, cols), dtype=_TORCH_OUT[np_out], device=cloud.device)
        with torch.cuda.device(cloud.device):
            rc = _lib.lib().nbr_tile_step_gather(
                mb.handle, ptr(cloud) if n else None, mb.code, n, edges_p, radii_p, len(radii),
                ptr(res) if res.numel() else None, int(res.shape[0]), out_code, 0,
                mb.boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                voxel_counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if voxel_counts is not None and n else None,
                offsets.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), stream_ptr(cloud.device))
        if rc != _lib.ERR_CAPACITY or attempt == 1:
            _lib.check(rc)
            break
        total = int(offsets[-1])
        mb.cache[key] = total
        if given and out_all.shape[0] < total:
            raise ValueError("out holds %d rows, the ranks produce %d" % (out_all.shape[0], total))
        mb.ensure_gather(int(_lib.lib().nbr_gather_staging_bytes(total, row_bytes)), group)
    total = int(offsets[-1])
    mb.cache[key] = total
    return res[:total]


def process_tile_host(cloud, edge_lengths, radii, out=None, out_dtype=np.float64, device=None, group=None, mailbox=None):
    """
    this rank's tile with HOST buffers: `cloud` (n_local, 3) numpy array or CPU tensor (float32 / float64; pinned
    memory is used in place), returns / fills a host array of (n_local, 4*S) rows (nbr_tile_step_host: the tile
    goes up, the rows come down in batches that overlap the kernels).  collective over the ranks.
    """
    from . import _lib, _results
    from .multiscale import _out_code
    assert len(edge_lengths) == len(radii), "edge_lengths and radii should be equal-length sequences."
    np_out, out_code = _out_code(out_dtype)
    arr = cloud.numpy() if isinstance(cloud, torch.Tensor) else np.asarray(cloud)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    arr = np.ascontiguousarray(arr)
    if arr.ndim != 2 or arr.shape[1] != 3:
        raise ValueError("wrong point cloud array shape")
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    tdt = torch.float32 if arr.dtype == np.float32 else torch.float64
    n = int(arr.shape[0])
    mb = mailbox if mailbox is not None else _mailboxes.get((id(group) if group is not None else 0, device.index, tdt))
    if mb is None:
        mb = HaloMailbox.connect_group(device, tdt, default_capacity(n), group)
        _mailboxes[(id(group) if group is not None else 0, device.index, tdt)] = mb
    if out is None:
        out = _results.empty((n, 4 * len(radii)), np_out)
    out_arr = out.numpy() if isinstance(out, torch.Tensor) else out
    if out_arr.shape != (n, 4 * len(radii)) or out_arr.dtype != np_out or not out_arr.flags["C_CONTIGUOUS"]:
        raise ValueError("out has the wrong shape, dtype or layout")
    edges_arr, edges_p = _lib.f64_array(list(edge_lengths))
    radii_arr, radii_p = _lib.f64_array(list(radii))
    with torch.cuda.device(device):
        _lib.check(_lib.lib().nbr_tile_step_host(
            mb.handle, ctypes.c_void_p(arr.ctypes.data) if n else None, mb.code, n, edges_p, radii_p, len(radii),
            ctypes.c_void_p(out_arr.ctypes.data) if n else None, out_code, 0,
            mb.boxes.ctypes.data_as(ctypes.POINTER(ctypes.c_double))))
    return out


def process_tiles_local(clouds, edge_lengths, radii, out_dtype=np.float32, mailboxes=None, capacity_rows=None, gather=False):
    """
    several tiles in ONE process (one device, or one tile per visible device): the same mailbox path as the
    multi-process run, every step issued for all tiles before the next one.  returns the list of per-tile
    feature tensors (rows in each tile's own order).  with one device this is also a way to process a cloud
    tile by tile; the single-GPU tests use it to prove tile + halo == unpartitioned.
    gather=True: every tile's rows also go into the staging buffer of every other tile from inside the feature kernel
    (the peer-store all-gather of process_tile(gather=True)); returns one (sum n, 4*S) tensor per tile, all equal.
    """
    from .multiscale import _out_code
    assert len(edge_lengths) == len(radii), "edge_lengths and radii should be equal-length sequences."
    np_out, out_code = _out_code(out_dtype)
    world = len(clouds)
    own = mailboxes is None
    if own:
        cap = capacity_rows if capacity_rows is not None else default_capacity(max(int(c.shape[0]) for c in clouds))
        mailboxes = HaloMailbox.local_set(world, [c.device for c in clouds], clouds[0].dtype, cap)
    try:
        for mb, c in zip(mailboxes, clouds):
            mb.publish(c)
        geo = []
        for mb, c in zip(mailboxes, clouds):
            boxes = mb.wait_boxes()
            geo.append(_tile_geometry(boxes, mb.rank, edge_lengths, radii))
        for mb, c, g in zip(mailboxes, clouds, geo):
            mb.push(c, g[0])
        outs = []
        if gather:
            from . import _lib
            from ._util import ptr, stream_ptr
            from .multiscale import _TORCH_OUT
            devices = [c.device for c in clouds]
            one_device = len(set(devices)) == 1
            if not one_device and len(set(devices)) != world:
                raise ValueError("gather=True needs all tiles on one device or every tile on its own")
            sizes = [int(c.shape[0]) for c in clouds]
            offs = np.ascontiguousarray(np.concatenate([[0], np.cumsum(sizes)]), dtype=np.int64)
            cols = 4 * len(radii)
            row_bytes = cols * np.dtype(np_out).itemsize
            HaloMailbox.ensure_gather_local(mailboxes, int(_lib.lib().nbr_gather_staging_bytes(int(offs[-1]), row_bytes)))
            f64p, i64p = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)
            edges_arr, edges_p = _lib.f64_array(list(edge_lengths))
            radii_arr, radii_p = _lib.f64_array(list(radii))
            outs = [torch.empty((int(offs[-1]), cols), dtype=_TORCH_OUT[np_out], device=c.device) for c in clouds]
            for mb, c, res, (h, glob, mine, local, origin) in zip(mailboxes, clouds, outs, geo):
                with torch.cuda.device(c.device):
                    stream = torch.cuda.current_stream(c.device)
                    perm, ordered = _order_tile(c, mine, origin, float(min(edge_lengths)), stream)
                    n = int(c.shape[0])
                    _lib.check(_lib.lib().nbr_multiscale_features_tile_mb_gather(
                        ptr(ordered) if n else None, ptr(perm) if n else None, mb.code, n, mb.handle, local.ctypes.data_as(f64p),
                        glob.ctypes.data_as(f64p), edges_p, radii_p, len(radii), ptr(res), out_code, 0, int(offs[mb.rank]),
                        int(offs[-1]), None, stream_ptr(c.device)))
            if not one_device:
                # one tile per device: signal + wait on every device (on ONE device the launches are stream-ordered and a
                # waiting kernel would only block the tiles behind it)
                for mb in mailboxes:
                    with torch.cuda.device(mb.device):
                        _lib.check(_lib.lib().nbr_gather_finish(mb.handle, stream_ptr(mb.device)))
            for mb, res in zip(mailboxes, outs):
                with torch.cuda.device(mb.device):
                    _lib.check(_lib.lib().nbr_gather_unpermute(mb.handle, offs.ctypes.data_as(i64p), row_bytes, ptr(res), stream_ptr(mb.device)))
        for mb, c, (h, glob, mine, local, origin) in zip(mailboxes, clouds, geo) if not gather else ():
            with torch.cuda.device(c.device):
                stream = torch.cuda.current_stream(c.device)
                perm, ordered = _order_tile(c, mine, origin, float(min(edge_lengths)), stream)
                out = _check_out(None, int(c.shape[0]), len(radii), np_out, c.device)
                _tile_features_mb(mb, ordered, perm, local, glob, edge_lengths, radii, out, out_code)
            outs.append(out)
        for mb in mailboxes:
            torch.cuda.synchronize(mb.device)
            st = mb.status()
            if st["timeout"] or st["dropped"]:
                raise RuntimeError("halo mailbox of tile %d: %s" % (mb.rank, st))
        return outs
    finally:
        if own:
            for mb in mailboxes:
                mb.close()


def _use_mailboxes(cloud):
    return (cloud.is_cuda and cloud.dtype in (torch.float32, torch.float64) and cloud.is_contiguous()
            and os.environ.get("NBR_HALO", "mailbox") != "nccl")


def gather_mode(gather, world, mailbox_path, env=None):
    """how process_tile gathers: None (no gather), "peer" (rows stored into the peers' staging buffers by the feature
    kernel) or "nccl" (all-gather after the kernel).  default: peer stores up to 4 ranks -- the receivers put (N - 1) n
    rows in place with random 80-byte writes (0.9 ms per 10M rows, DRAM-bound); from 8 ranks on that costs what riding on
    the kernel saves.  NBR_GATHER=peer|nccl|copy overrides the default, gather="peer"|"nccl" overrides both."""
    if not gather:
        return None
    if gather == "peer":
        if not mailbox_path:
            raise ValueError('gather="peer" needs the CUDA mailbox path')
        return "peer"
    if gather == "nccl" or not mailbox_path:
        return "nccl"
    env = os.environ.get("NBR_GATHER") if env is None else env
    auto = env if env else ("peer" if world <= 4 else "nccl")
    return "peer" if auto in ("peer", "fused", "copy") else "nccl"


def process_tile(cloud, edge_lengths, radii, out=None, out_dtype=np.float32, gather=False, group=None,
                 compute=None, voxel_counts=None, out_all=None):
    """
    features of this rank's tile (n_local, 4*S).  `cloud`: (n_local, 3) tensor on this rank's device; a rank may
    hold an empty tile (it still takes part in every collective step).
    gather=True: returns the rows of every rank, concatenated in rank order, on every rank.  on the CUDA mailbox path
    the all-gather happens without a collective call: the fused feature kernel stores every finished row into a
    staging buffer of every other rank over NVLink while it computes, the receivers put the rows in place
    (nbr_tile_step_gather; the default up to 4 ranks, gather="peer" / NBR_GATHER=peer forces it; gather="nccl" /
    NBR_GATHER=nccl: features, then an NCCL all-gather).  out_all: optional preallocated (>= sum n, 4*S) tensor for
    the rows of all ranks (peer-store path).
    compute(query, search, edges, radii, (lo, hi) numpy, out_dtype, out) -> features; default = CUDA path.
    voxel_counts: optional int64 numpy array (S,) receiving the unique voxels per scale of this rank's lattices
    (tile + halo; mailbox path only, forces a synchronisation).
    """
    assert len(edge_lengths) == len(radii), "edge_lengths and radii should be equal-length sequences."
    world = dist.get_world_size(group)
    sizes = None
    mailbox_path = compute is None and _use_mailboxes(cloud) and world <= 16
    if gather_mode(gather, world, mailbox_path) == "peer":
        res = _process_tile_gather(cloud, edge_lengths, radii, out_all, out_dtype, group, voxel_counts)
        if out is not None:
            # the rows of this rank's tile are a slice of the gathered result; a caller that also wants them in `out`
            # gets a copy
            n = int(cloud.shape[0])
            first = int(sum(int(v) for v in _group_mailbox(cloud, group).boxes[:dist.get_rank(group), 6]))
            if tuple(out.shape) != (n, res.shape[1]) or out.dtype != res.dtype or out.device != res.device:
                raise ValueError("out has the wrong shape, dtype or device")
            out.copy_(res[first:first + n])
        return res
    if compute is None and _use_mailboxes(cloud) and world <= 16:
        feats, boxes = _process_tile_cuda(cloud, edge_lengths, radii, out, out_dtype, group, voxel_counts)
        sizes = [int(v) for v in boxes[:, 6]]
    else:
        halo, (g_lo, g_hi), _ = exchange_halo(cloud, edge_lengths, radii, group)
        search = torch.cat([cloud, halo], 0) if halo.numel() else cloud
        bbox = (g_lo.cpu().numpy(), g_hi.cpu().numpy())
        # the queries are handed over as the first rows of the search buffer: the CUDA path then orders tile + halo
        # once, builds the lattices from the ordered copy and keeps only the tile's points as queries
        query = search[:cloud.shape[0]]
        if cloud.shape[0] == 0:
            from .multiscale import _TORCH_OUT
            feats = torch.zeros((0, 4 * len(radii)), dtype=_TORCH_OUT[np.dtype(out_dtype).type], device=cloud.device)
        else:
            feats = (compute or _gpu_compute)(query, search, edge_lengths, radii, bbox, out_dtype, out)
    if not gather:
        return feats
    return gather_rows(feats, sizes, group)


def gather_rows(feats, sizes=None, group=None):
    """all-gather-v of the feature rows into ONE preallocated (sum n, C) tensor, rank order.  sizes: every
    rank's row count if the caller already has it (the mailbox path carries it in the box table)."""
    world = dist.get_world_size(group)
    feats = feats.contiguous()
    if sizes is None:
        n_local = torch.tensor([feats.shape[0]], dtype=torch.int64, device=feats.device)
        got = [torch.empty_like(n_local) for _ in range(world)]
        dist.all_gather(got, n_local, group=group)
        sizes = [int(v.item()) for v in got]
    total = torch.empty((sum(sizes), feats.shape[1]), dtype=feats.dtype, device=feats.device)
    if len(set(sizes)) == 1 and feats.is_cuda:
        dist.all_gather_into_tensor(total, feats, group=group)        # equal tiles: straight into the result
    elif not feats.is_cuda:
        # gloo has no uneven all-gather: pad every contribution to the largest tile
        cap = max(sizes)
        padded = feats if feats.shape[0] == cap else torch.cat([feats, feats.new_zeros((cap - feats.shape[0], feats.shape[1]))], 0)
        parts = [torch.empty((cap, feats.shape[1]), dtype=feats.dtype) for _ in range(world)]
        dist.all_gather(parts, padded.contiguous(), group=group)
        offs = np.concatenate([[0], np.cumsum(sizes)])
        for r in range(world):
            total[int(offs[r]):int(offs[r + 1])] = parts[r][:sizes[r]]
    else:
        offs = np.concatenate([[0], np.cumsum(sizes)])
        views = [total[int(offs[r]):int(offs[r + 1])] for r in range(world)]
        dist.all_gather(views, feats, group=group)                     # uneven: one broadcast per rank into its rows
    return total
