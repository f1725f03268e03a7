"""torch.autograd.Function wrappers over the C ABI (include/coverage_b200.h).

PyTorch is plumbing here: it owns device memory, streams and (when the cloud is sharded over
ranks) the NCCL communicator.  Every numeric result comes from libcovb200.so.

Sharded use: each rank passes ITS slice of the cloud plus a `group`; the per-pose normalisers
are all-reduced with MIN/MAX, the accumulators with SUM (a few KB), so every rank ends up with
the same scalar and the same pose gradients, while per-point outputs stay sharded.
"""
import contextlib
import ctypes
import os
import threading

import torch

from . import _lib


def _stream(device=None):
    """The caller's current stream ON THE TENSORS' DEVICE (not on whatever device happens to be current)."""
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _guard(t):
    """Device guard for a C-ABI call: the library launches on the CUDA device that is current, so the tensors'
    device is made current for the duration of the call (`ModelPose(..., device='cuda:1')` while cuda:0 is current)."""
    return torch.cuda.device(t.device)


def _call(name, on, *args):
    """One C-ABI call on the device of tensor `on` and on that device's current stream (every entry point takes the
    stream last); raises RuntimeError with the library's message on a non-zero return code.  The device guard is only
    entered when the tensor's device is not the current one (the common case pays two cheap queries)."""
    fn = getattr(_lib.lib(), name)
    dev = on.device
    if dev.index == torch.cuda.current_device():
        rc = fn(*args, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    else:
        with torch.cuda.device(dev):
            rc = fn(*args, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc:
        _lib.check(rc, name)


def _dev_f32(t, device=None, what="tensor"):
    """fp32, contiguous, on a CUDA device, 16-byte aligned (clones only when it has to)."""
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{what} must be a torch.Tensor")
    if device is not None and t.device != device:
        t = t.to(device)
    if not t.is_cuda:
        raise RuntimeError(f"{what} is on {t.device}: the coverage ops are CUDA-only (no CPU fallback)")
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    if not t.is_contiguous():
        t = t.contiguous()
    if t.data_ptr() % 16:
        t = t.clone()
    return t


class _Mode(threading.local):
    dense = False      # evaluate every (point, pose) pair (cov_traj_opts.dense)
    stats = None       # CUDA int64 tensor of 8 work counters, or None (cov_traj_opts.stats_dev)


_MODE = _Mode()


@contextlib.contextmanager
def evaluation(dense=None, stats=None):
    """Per-thread defaults of the trajectory ops inside the block: `dense=True` switches the exact pruning off
    (A/B measurements, unordered clouds); `stats` = a CUDA int64 tensor of 8 counters the kernels add to.  Python-side
    only: the C ABI takes these per call (cov_traj_opts) and keeps no process-wide switch."""
    old = (_MODE.dense, _MODE.stats)
    if dense is not None:
        _MODE.dense = bool(dense)
    if stats is not None:
        _MODE.stats = stats
    try:
        yield
    finally:
        _MODE.dense, _MODE.stats = old


def _opts(dense=None, rewards_prefilled=False, prefill=None):
    return _lib.traj_opts(_MODE.dense if dense is None else dense, rewards_prefilled, _MODE.stats, prefill)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


def _all_reduce(t, op, group):
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(t, op=op, group=group)


def _reduce_ops():
    import torch.distributed as dist
    return dist.ReduceOp.MIN, dist.ReduceOp.MAX, dist.ReduceOp.SUM


class PeerExchange:
    """The world's exchange buffers for the in-kernel NVLink all-reduces of one (group, device, pose count): symmetric
    memory allocated and rendezvoused ONCE (outside any graph capture), then used by `cov_peer_allreduce`
    (csrc/cov_peer.cu) for the 2 W normalisers and the 22 W + 1 accumulator doubles of every step."""

    def __init__(self, group, device, W):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        L = _lib.lib()
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f"peer exchange supports up to {_lib.MAX_PEERS} ranks")
        self.n_mm, self.n_acc = 2 * W, W * _lib.ACC_STRIDE + 1
        b_mm = (L.cov_peer_region_bytes(_lib.PEER_MINMAX_F32, self.n_mm, self.world) + 255) // 256 * 256
        b_acc = (L.cov_peer_region_bytes(_lib.PEER_SUM_F64, self.n_acc, self.world) + 255) // 256 * 256
        self.off_mm, self.off_acc = 0, b_mm
        with torch.cuda.device(device):
            self.buf = symm.empty(b_mm + b_acc, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self.handle = symm.rendezvous(self.buf, group)
            torch.cuda.synchronize(device)
        dist.barrier(group)      # every rank's buffer is zeroed before anybody pushes into it
        self.peers = _lib.Peers()
        for r, ptr in enumerate(self.handle.buffer_ptrs):
            self.peers.ptr[r] = ptr
        self.peers.world, self.peers.rank = self.world, self.rank

    def reduce(self, kind, t, offset):
        _call("cov_peer_allreduce", t, kind, _ptr(t), t.numel(), ctypes.byref(self.peers), offset)


_PEER_CACHE = {}
_PEER_STATE = {"enabled": os.environ.get("COV_PEER_EXCHANGE", "1") != "0", "warned": False}


def peer_exchange(group, device, W):
    """The PeerExchange for (group, device, W), or None when the in-kernel exchange is off or unavailable (then the
    collectives are NCCL all-reduces).  Created on first use, which must happen outside CUDA-graph capture (any eager
    warm-up step does it)."""
    if group is None or not _PEER_STATE["enabled"] or device.type != "cuda" or W > 2048:
        return None
    key = (id(group), device.index, W)
    if key not in _PEER_CACHE:
        try:
            import torch.distributed as dist
            if dist.get_backend(group) != "nccl":
                raise RuntimeError("peer exchange needs one CUDA device per rank (nccl group)")
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("first use inside a CUDA-graph capture")
            _PEER_CACHE[key] = PeerExchange(group, device, W)
        except Exception as exc:   # no P2P / symmetric memory on this box: the NCCL path is equivalent
            _PEER_CACHE[key] = None
            if not _PEER_STATE["warned"]:
                _PEER_STATE["warned"] = True
                import warnings
                warnings.warn(f"in-kernel NVLink exchange unavailable ({str(exc).splitlines()[0][:160]}); using NCCL all-reduces")
    return _PEER_CACHE[key]


def _all_reduce_minmax(minmax, W, group):
    """Global per-pose minima (first W) and maxima (last W) in ONE exchange: the in-kernel NVLink all-reduce
    (MIN | MAX halves) when available, else one NCCL MAX with the minima negated (MIN(x) = -MAX(-x), exact)."""
    if group is None:
        return
    px = peer_exchange(group, minmax.device, W) if minmax.is_cuda and minmax.dtype == torch.float32 else None
    if px is not None and px.n_mm == minmax.numel():
        px.reduce(_lib.PEER_MINMAX_F32, minmax, px.off_mm)
        return
    minmax[:W].neg_()
    _all_reduce(minmax, _reduce_ops()[1], group)
    minmax[:W].neg_()


def _all_reduce_acc(acc, W, group):
    """SUM of the 22 W + 1 accumulator doubles over the ranks (in rank order: identical on every rank)."""
    if group is None:
        return
    px = peer_exchange(group, acc.device, W) if acc.is_cuda and acc.dtype == torch.float64 else None
    if px is not None and px.n_acc == acc.numel():
        px.reduce(_lib.PEER_SUM_F64, acc, px.off_acc)
        return
    _all_reduce(acc, _reduce_ops()[2], group)


class CudaBackend:
    """The five C-ABI calls of the coverage path on device tensors.  `ops._BACKEND` is the only instance the
    product uses; the N>1 CPU tests swap in a stand-in with the same methods to exercise the collective
    plumbing below over gloo."""

    def prepare(self, t, device=None, what="tensor"):
        return _dev_f32(t, device, what)

    def order_cloud(self, pts, spatial_sort_cloud=True):
        """Once per cloud (ModelTraj): Morton-ordered copy, its permutation and the 128-point boxes."""
        perm = None
        if spatial_sort_cloud and pts.shape[0] > 0:
            pts, perm = spatial_sort(pts)
        return pts, perm, tile_boxes(pts)

    def pose_fused(self, pts, t, q, Kd, cam, w, obs):
        L = _lib.lib()
        n = pts.shape[0]
        acc = torch.empty(_lib.POSE_ACC, dtype=torch.float64, device=pts.device)
        ws_bytes = L.cov_pose_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pts.device)
        _call("cov_pose_fused", pts, _ptr(pts), n, _ptr(w), _ptr(t), _ptr(q), _ptr(Kd), ctypes.byref(cam), _ptr(obs),
                                    _ptr(acc), _ptr(ws), ws_bytes)
        return acc

    def pose_epilogue(self, acc, t, q):
        out = torch.empty(8, dtype=torch.float32, device=acc.device)
        _call("cov_pose_epilogue", acc, _ptr(acc), _ptr(t), _ptr(q), _ptr(out))
        return out

    def traj_workspace_bytes(self, pts, W):
        return _lib.lib().cov_traj_workspace_bytes(pts.shape[0], W)

    def traj_workspace(self, pts, W):
        return torch.empty(self.traj_workspace_bytes(pts, W), dtype=torch.uint8, device=pts.device)

    def prefill_applies(self, pts, dense=None):
        """Will pass A on this cloud honour a `prefill` buffer (pruned path)?"""
        opts = _opts(dense)
        return bool(_lib.lib().cov_traj_prefill_applies(pts.shape[0], ctypes.byref(opts)))

    def traj_minmax(self, pts, P, Q, Kd, cam, boxes=None, ws=None, dense=None, prefill=None):
        W = P.shape[0]
        minmax = torch.empty(2 * W, dtype=torch.float32, device=pts.device)
        ws = self.traj_workspace(pts, W) if ws is None else ws
        opts = _opts(dense, prefill=prefill)
        _call("cov_traj_minmax", pts, _ptr(pts), pts.shape[0], _ptr(P), _ptr(Q), W, _ptr(Kd), ctypes.byref(cam),
              _ptr(boxes), _ptr(minmax), ctypes.byref(opts), _ptr(ws), ws.numel())
        return minmax

    def traj_fused(self, pts, P, Q, Kd, cam, minmax, upstream, rewards, reward_index=None, boxes=None, ws=None,
                   dense=None, prefilled=False):
        W, n = P.shape[0], pts.shape[0]
        acc = torch.empty(W * _lib.ACC_STRIDE + 1, dtype=torch.float64, device=pts.device)
        ws = self.traj_workspace(pts, W) if ws is None else ws
        opts = _opts(dense, rewards_prefilled=prefilled)
        _call("cov_traj_fused", pts, _ptr(pts), n, _ptr(P), _ptr(Q), W, _ptr(Kd), ctypes.byref(cam), _ptr(boxes),
              _ptr(minmax), _ptr(upstream), _ptr(reward_index), _ptr(rewards), _ptr(acc), ctypes.byref(opts), _ptr(ws),
              ws.numel())
        return acc

    def sweep_minmax(self, pts, P, Q, Kd, cam, boxes=None):
        """Per-pose minima and maxima of all T*P poses of a sweep (pass A, one pose-table-full at a time)."""
        L = _lib.lib()
        W, n, dev = P.shape[0], pts.shape[0], pts.device
        minmax = torch.empty(2 * W, dtype=torch.float32, device=dev)
        chunk = L.cov_traj_max_poses_pruned() if self.prefill_applies(pts) else L.cov_traj_max_poses()
        ws_bytes = L.cov_traj_workspace_bytes(n, min(W, chunk))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        opts = _opts()
        for w0 in range(0, W, chunk):
            w1 = min(W, w0 + chunk)
            mm = torch.empty(2 * (w1 - w0), dtype=torch.float32, device=dev)
            _call("cov_traj_minmax", pts, _ptr(pts), n, _ptr(P[w0:w1]), _ptr(Q[w0:w1]), w1 - w0, _ptr(Kd),
                  ctypes.byref(cam), _ptr(boxes), _ptr(mm), ctypes.byref(opts), _ptr(ws), ws_bytes)
            minmax[w0:w1] = mm[:w1 - w0]
            minmax[W + w0:W + w1] = mm[w1 - w0:]
        return minmax

    def sweep_sums(self, pts, P, Q, T, Pn, Kd, cam, boxes, minmax):
        """sum_j rewards_j(t) over this cloud for every trajectory t (fp64)."""
        L = _lib.lib()
        n, dev = pts.shape[0], pts.device
        ws_bytes = L.cov_sweep_workspace_bytes(n, T, Pn)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        sums = torch.zeros(T, dtype=torch.float64, device=dev)
        opts = _opts()
        _call("cov_sweep_rewards", pts, _ptr(pts), n, _ptr(P), _ptr(Q), T, Pn, _ptr(Kd), ctypes.byref(cam), _ptr(boxes),
              _ptr(minmax), _ptr(sums), ctypes.byref(opts), _ptr(ws), ws_bytes)
        return sums

    def traj_epilogue(self, acc, minmax, Q, n_total, upstream_mode):
        W = Q.shape[0]
        out = torch.empty(1 + 7 * W, dtype=torch.float32, device=acc.device)
        _call("cov_traj_epilogue", acc, _ptr(acc), _ptr(minmax), _ptr(Q), W, n_total, upstream_mode, _ptr(out))
        return out


_BACKEND = CudaBackend()


class CoveragePoseFn(torch.autograd.Function):
    """obs_j = dist_mask*fov_mask[*weight_j], total = sum_j obs_j (reference src/model.py:98-127)."""

    @staticmethod
    def forward(ctx, points, trans, quat, K, cam, weight, group):
        B = _BACKEND
        dev = points.device
        pts = B.prepare(points, what="points")
        t = B.prepare(trans, dev, "trans").reshape(3)
        q = B.prepare(quat, dev, "quat").reshape(4)
        Kd = B.prepare(K, dev, "intrins").reshape(9)
        w = None if weight is None else B.prepare(weight, dev, "weight").reshape(-1)
        n = pts.shape[0]
        obs = torch.empty(n, dtype=torch.float32, device=dev)
        out = CoveragePoseFn._run(pts, t, q, Kd, cam, w, obs, group)
        ctx.cam, ctx.group = cam, group
        ctx.shapes = (trans.shape, quat.shape)
        ctx.save_for_backward(pts, t, q, Kd, w, out)
        ctx.set_materialize_grads(False)
        return obs, out[0].clone()

    @staticmethod
    def _run(pts, t, q, Kd, cam, w, obs, group):
        acc = _BACKEND.pose_fused(pts, t, q, Kd, cam, w, obs)   # this shard's sums (8 doubles)
        if group is not None:
            _all_reduce(acc, _reduce_ops()[2], group)
        return _BACKEND.pose_epilogue(acc, t, q)

    @staticmethod
    def backward(ctx, g_obs, g_total):
        pts, t, q, Kd, w, out = ctx.saved_tensors
        g_t = g_q = None
        if g_total is not None:
            g_t = g_total * out[1:4]
            g_q = g_total * out[4:8]
        if g_obs is not None:
            # someone differentiated through the per-point vector: one more pass with
            # weight_j = upstream_j [* weight_j]
            up = _BACKEND.prepare(g_obs, pts.device, "grad_obs").reshape(-1)
            if w is not None:
                up = up * w
            o2 = CoveragePoseFn._run(pts, t, q, Kd, ctx.cam, up, None, ctx.group)
            g_t = o2[1:4] if g_t is None else g_t + o2[1:4]
            g_q = o2[4:8] if g_q is None else g_q + o2[4:8]
        ts, qs = ctx.shapes
        return (None, None if g_t is None else g_t.reshape(ts), None if g_q is None else g_q.reshape(qs),
                None, None, None, None)


class CoverageTrajFn(torch.autograd.Function):
    """rewards_j = sigmoid(sum_w logit(clip(normalised m_jw))), mean = mean_j rewards_j
    over the poses given (reference src/model.py:217-237, :246)."""

    @staticmethod
    def forward(ctx, points, poses, quats, K, cam, n_total, group, reward_index=None, boxes=None, dense=None,
                workspace=None):
        B = _BACKEND
        dev = points.device
        pts = B.prepare(points, what="points")
        if reward_index is not None and (reward_index.dtype != torch.int32 or reward_index.shape[0] != pts.shape[0]
                                         or reward_index.device != dev or not reward_index.is_contiguous()):
            raise ValueError("reward_index must be a contiguous int32 tensor with one entry per point, on the cloud's device")
        P = B.prepare(poses, dev, "poses").reshape(-1, 3)
        Q = B.prepare(quats, dev, "quats").reshape(-1, 4)
        Kd = B.prepare(K, dev, "intrins").reshape(9)
        W, n = P.shape[0], pts.shape[0]
        if Q.shape[0] != W:
            raise ValueError("poses and quats disagree on the number of waypoints")
        n_total = int(n if n_total is None else n_total)
        dense = _MODE.dense if dense is None else bool(dense)
        ws = workspace                                       # shared by both passes; a caller may keep one per cloud
        if ws is None or ws.numel() < B.traj_workspace_bytes(pts, W) or ws.device != dev:
            ws = B.traj_workspace(pts, W)
        rewards = torch.empty(n, dtype=torch.float32, device=dev)
        # the pruned pass A pre-fills pass B's rewards with 1/2 under its idle memory bandwidth
        prefill = rewards if B.prefill_applies(pts, dense) else None
        minmax = B.traj_minmax(pts, P, Q, Kd, cam, boxes, ws, dense, prefill)   # pass A on this shard
        _all_reduce_minmax(minmax, W, group)                 # global normalisers: W minima, W maxima
        out = CoverageTrajFn._run(pts, P, Q, Kd, cam, minmax, None, rewards, n_total, group, reward_index, boxes, ws,
                                  dense, prefill is not None)
        ctx.cam, ctx.group, ctx.n_total, ctx.reward_index, ctx.boxes = cam, group, n_total, reward_index, boxes
        ctx.dense = dense
        ctx.shapes = (poses.shape, quats.shape)
        ctx.save_for_backward(pts, P, Q, Kd, minmax, out)
        ctx.set_materialize_grads(False)
        return rewards, out[0].clone()

    @staticmethod
    def _run(pts, P, Q, Kd, cam, minmax, upstream, rewards, n_total, group, reward_index=None, boxes=None, ws=None,
             dense=None, prefilled=False):
        acc = _BACKEND.traj_fused(pts, P, Q, Kd, cam, minmax, upstream, rewards, reward_index, boxes, ws, dense,
                                  prefilled)   # pass B
        _all_reduce_acc(acc, P.shape[0], group)
        return _BACKEND.traj_epilogue(acc, minmax, Q, n_total, 0 if upstream is None else 1)

    @staticmethod
    def backward(ctx, g_rewards, g_mean):
        pts, P, Q, Kd, minmax, out = ctx.saved_tensors
        W = P.shape[0]
        g_p = g_q = None
        if g_mean is not None:
            g_p = g_mean * out[1:1 + 3 * W]
            g_q = g_mean * out[1 + 3 * W:]
        if g_rewards is not None:
            up = _BACKEND.prepare(g_rewards, pts.device, "grad_rewards").reshape(-1)
            scratch = torch.empty(pts.shape[0], dtype=torch.float32, device=pts.device)
            o2 = CoverageTrajFn._run(pts, P, Q, Kd, ctx.cam, minmax, up, scratch, ctx.n_total, ctx.group,
                                     ctx.reward_index, ctx.boxes, None, ctx.dense)
            g_p = o2[1:1 + 3 * W] if g_p is None else g_p + o2[1:1 + 3 * W]
            g_q = o2[1 + 3 * W:] if g_q is None else g_q + o2[1 + 3 * W:]
        ps, qs = ctx.shapes
        return (None, None if g_p is None else g_p.reshape(ps), None if g_q is None else g_q.reshape(qs),
                None, None, None, None, None, None, None, None)


class TrajRegularizersFn(torch.autograd.Function):
    """(l2, smooth, length) of ModelTraj.criterion (reference src/model.py:248-259) and their gradients w.r.t. the
    waypoints in one launch (cov_traj_regularizers)."""

    @staticmethod
    def forward(ctx, poses, poses0, smoothness_weight, traj_length_weight, eps):
        P = _dev_f32(poses, what="poses").reshape(-1, 3)
        P0 = _dev_f32(poses0, P.device, "poses0").reshape(-1, 3)
        W = P.shape[0]
        out = torch.empty(3 + 9 * W, dtype=torch.float32, device=P.device)
        _call("cov_traj_regularizers", P, _ptr(P), _ptr(P0), W, float(smoothness_weight), float(traj_length_weight),
                                                    float(eps), _ptr(out))
        ctx.save_for_backward(out)
        ctx.shape = poses.shape
        return out[:3].clone()

    @staticmethod
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        W = (out.numel() - 3) // 9
        grads = out[3:].reshape(3, W * 3)
        return (g.reshape(1, 3) @ grads).reshape(ctx.shape), None, None, None, None


def traj_regularizers(poses, poses0, smoothness_weight, traj_length_weight, eps=1e-6):
    """Returns a (3,) tensor (l2, smooth, length), differentiable w.r.t. `poses` ((W,3), W >= 3, CUDA)."""
    return TrajRegularizersFn.apply(poses, poses0, smoothness_weight, traj_length_weight, eps)


def coverage_pose(points, trans, quat, intrins, img_width, img_height, min_dist=1.0, max_dist=5.0, eps=1e-6,
                  weight=None, group=None):
    """Fused ModelPose objective.  Returns (observations (N,), total = sum(observations));
    differentiable w.r.t. trans (.., 3) and quat (.., 4) (w, x, y, z)."""
    cam = _lib.camera(img_width, img_height, min_dist, max_dist, eps)
    return CoveragePoseFn.apply(points, trans, quat, intrins, cam, weight, group)


def coverage_traj(points, poses, quats, intrins, img_width, img_height, min_dist=1.0, max_dist=5.0, eps=1e-6,
                  n_total=None, group=None, reward_index=None, boxes=None, dense=None, workspace=None):
    """Fused ModelTraj visibility term over the W poses given.  Returns (rewards (N,), mean(rewards)).
    `reward_index` (int32, from `spatial_sort`): `points` is a reordered copy of the caller's cloud and
    rewards come back in the caller's order.  `boxes` (from `tile_boxes(points)`): built once per cloud so the
    pruned kernels do not rebuild them on every call.  `dense=True`: evaluate every pair (no pruning; what an unordered
    cloud should ask for).  `workspace`: a uint8 CUDA tensor kept by the caller across calls (one per cloud)."""
    cam = _lib.camera(img_width, img_height, min_dist, max_dist, eps)
    return CoverageTrajFn.apply(points, poses, quats, intrins, cam, n_total, group, reward_index, boxes, dense, workspace)


@torch.no_grad()
def tile_boxes(points):
    """Bounding boxes of runs of 128 consecutive points (include/coverage_b200.h: cov_tile_boxes), (nb, 8) fp32."""
    L = _lib.lib()
    pts = _dev_f32(points, what="points")
    n = pts.shape[0]
    boxes = torch.empty(max(int(L.cov_tile_boxes_count(n)), 1), 8, dtype=torch.float32, device=pts.device)
    if n > 0:
        _call("cov_tile_boxes", pts, _ptr(pts), n, _ptr(boxes))
    return boxes


@torch.no_grad()
def spatial_sort(points):
    """Morton-order copy of a cloud, made once per cloud (see include/coverage_b200.h: cov_spatial_sort).
    Returns (sorted (N,3) fp32, perm (N,) int32) with sorted[j] = points[perm[j]]."""
    L = _lib.lib()
    pts = _dev_f32(points, what="points")
    n = pts.shape[0]
    out = torch.empty_like(pts)
    perm = torch.empty(n, dtype=torch.int32, device=pts.device)
    if n == 0:
        return out, perm
    ws_bytes = L.cov_spatial_sort_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=pts.device)
    _call("cov_spatial_sort", pts, _ptr(pts), n, _ptr(out), _ptr(perm), _ptr(ws), ws_bytes)
    return out, perm


@torch.no_grad()
def sort_pairs(keys, vals, begin_bit=0, end_bit=32):
    """Stable radix sort of (keys int32/uint32 bit patterns, vals int32) CUDA tensors by key bits [begin_bit, end_bit)
    (include/coverage_b200.h: cov_sort_pairs); returns sorted copies.  Keys compare as UNSIGNED 32-bit values."""
    L = _lib.lib()
    if not (keys.is_cuda and vals.is_cuda and keys.dtype == torch.int32 and vals.dtype == torch.int32 and keys.shape == vals.shape
            and keys.dim() == 1):
        raise RuntimeError("sort_pairs: keys and vals must be 1-D int32 CUDA tensors of equal length")
    k, v = keys.clone(), vals.clone()
    n = k.shape[0]
    if n == 0:
        return k, v
    ws_bytes = L.cov_sort_pairs_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=k.device)
    _call("cov_sort_pairs", k, _ptr(k), _ptr(v), n, int(begin_bit), int(end_bit), _ptr(ws), ws_bytes)
    return k, v


@torch.no_grad()
def sweep_rewards(points, poses, quats, intrins, img_width, img_height, min_dist=1.0, max_dist=5.0, eps=1e-6,
                  n_total=None, group=None, boxes=None, presorted=False, shard="points"):
    """Forward-only mean reward of many candidate trajectories: poses (T, P, 3), quats (T, P, 4) -> (T,) fp64
    (semantics: ModelTraj.forward per trajectory with every pose evaluated, reference src/model.py:217-237).
    The cloud is Morton-ordered first (no per-point output exists, so the order is free) unless `presorted`;
    callers that sweep the same cloud repeatedly order it once (`spatial_sort`, `tile_boxes`) and pass both.

    With a `group` (one process per GPU) there are two ways to split the work:
      shard="points"        `points` is THIS RANK'S slice of the cloud; the per-pose normalisers are all-reduced (MAX of
                            2 T P floats) and the per-trajectory sums are all-reduced (SUM of T doubles);
      shard="trajectories"  every rank holds the WHOLE cloud (it fits: 50 M points = 600 MB) and evaluates its own
                            contiguous block of T / world trajectories — no collective on the data path, one all-gather
                            of T doubles at the end.  The choice when the cloud fits one GPU."""
    B = _BACKEND
    if shard not in ("points", "trajectories"):
        raise ValueError("shard must be 'points' or 'trajectories'")
    pts = B.prepare(points, what="points")
    if not presorted and boxes is None and pts.shape[0] > 0:
        pts, _, boxes = B.order_cloud(pts, True)
    dev = pts.device
    T, Pn = poses.shape[0], poses.shape[1]
    n = pts.shape[0]
    Kd = B.prepare(intrins, dev, "intrins").reshape(9)
    cam = _lib.camera(img_width, img_height, min_dist, max_dist, eps)
    if group is not None and shard == "trajectories":
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        per = (T + world - 1) // world
        t0, t1 = min(T, rank * per), min(T, (rank + 1) * per)
        mine = torch.zeros(per, dtype=torch.float64, device=dev)
        if t1 > t0:
            P = B.prepare(poses[t0:t1], dev, "poses").reshape(-1, 3)
            Q = B.prepare(quats[t0:t1], dev, "quats").reshape(-1, 4)
            minmax = B.sweep_minmax(pts, P, Q, Kd, cam, boxes)
            mine[:t1 - t0] = B.sweep_sums(pts, P, Q, t1 - t0, Pn, Kd, cam, boxes, minmax)
        out = torch.empty(world * per, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, mine, group=group)
        return out[:T] / float(n)
    P = B.prepare(poses, dev, "poses").reshape(-1, 3)
    Q = B.prepare(quats, dev, "quats").reshape(-1, 4)
    minmax = B.sweep_minmax(pts, P, Q, Kd, cam, boxes)
    _all_reduce_minmax(minmax, T * Pn, group)
    sums = B.sweep_sums(pts, P, Q, T, Pn, Kd, cam, boxes, minmax)
    if group is not None:
        _all_reduce(sums, _reduce_ops()[2], group)
    return sums / float(n if n_total is None else n_total)


@torch.no_grad()
def frustum_cull(points_nx3, intrins, img_width, img_height, min_dist=1.0, max_dist=10.0):
    """Binary frustum test + ordered compaction (reference src/tools.py:176-187).
    Returns (idx int64 (M,), dist_mask bool (N,), fov_mask bool (N,))."""
    L = _lib.lib()
    pts = _dev_f32(points_nx3, what="points")
    dev = pts.device
    Kd = _dev_f32(intrins, dev, "intrins")[:3, :3].contiguous().reshape(9)
    n = pts.shape[0]
    dm = torch.empty(n, dtype=torch.uint8, device=dev)
    fm = torch.empty(n, dtype=torch.uint8, device=dev)
    idx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = L.cov_cull_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _call("cov_frustum_cull", pts, _ptr(pts), n, _ptr(Kd), float(img_width), float(img_height), float(min_dist),
                                  float(max_dist), _ptr(dm), _ptr(fm), _ptr(idx), _ptr(cnt), _ptr(ws), ws_bytes)
    m = int(cnt.item())
    return idx[:m].long(), dm.view(torch.bool), fm.view(torch.bool)   # the masks hold 0/1 bytes: zero-copy views


@torch.no_grad()
def spherical_flip(points, param):
    """Bit-exact fp32 spherical flip (reference src/tools.py:38-53). Returns (flipped (N,3), radius 0-dim)."""
    L = _lib.lib()
    pts = _dev_f32(points, what="points")
    n = pts.shape[0]
    out = torch.empty_like(pts)
    rad = torch.empty(2, dtype=torch.float32, device=pts.device)
    _call("cov_hpr_flip", pts, _ptr(pts), n, float(10.0 ** param), _ptr(out), _ptr(rad))
    return out, rad[0]


@torch.no_grad()
def hpr_hull_mask(flipped, return_info=False):
    """Vertex mask of conv(flipped U {origin}) (reference src/tools.py:56-64 + the vertex set of :79).
    Returns (mask uint8 (N,), origin_is_vertex bool, n_exact_fallback int) [+ the raw 4-int info when `return_info`]."""
    L = _lib.lib()
    f = _dev_f32(flipped, what="flipped points")
    n = f.shape[0]
    mask = torch.empty(n, dtype=torch.uint8, device=f.device)
    info = torch.zeros(4, dtype=torch.int32, device=f.device)
    ws_bytes = L.cov_hpr_hull_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=f.device)
    _call("cov_hpr_hull", f, _ptr(f), n, _ptr(mask), _ptr(info), _ptr(ws), ws_bytes)
    info_h = info.tolist()
    if return_info:
        return mask, bool(info_h[0]), int(info_h[1]), info_h
    return mask, bool(info_h[0]), int(info_h[1])
