"""Symmetric InfoNCE head, local or sharded over a data-parallel group.

The reference has no distributed code; the sharded scheme is defined by the north star:
each rank encodes its batch shard, the L2-normalised embeddings are all-gathered (NCCL over NVLink),
every rank scores its speech rows and its EEG columns against the global batch, the three per-row
vectors (row LSE, column LSE, diagonal) are all-gathered, and the loss is identical on every rank.
The backward recomputes the two logit blocks, so no embedding-gradient reduce-scatter is needed;
parameter gradients are all-reduced with SUM (each rank already holds exact global-loss gradients
with respect to its own rows).  With world size 1 this is exactly clip_model.py:675-693.

Device compute goes through the C ABI (``CudaHeadOps``).  The collective plumbing is written against
a small ops interface so that the world_size-2 gloo tests can drive the very same code on CPU with a
checker implementation injected by the test-suite; the product path never leaves ``CudaHeadOps``.
"""
import ctypes

import torch
import torch.distributed as dist

from . import _lib as L


class CudaHeadOps:
    """Head kernels behind include/eegclip.h (eegclip_l2norm_*, eegclip_infonce_*)."""

    @staticmethod
    def _scratch(b, Bg, D, device):
        n = ctypes.c_size_t()
        L.call("eegclip_infonce_workspace", b, Bg, D, ctypes.byref(n))
        return torch.empty(max(n.value, 16), dtype=torch.uint8, device=device)

    def l2norm_fwd(self, x):
        x = L.f32c(x)
        xn = torch.empty_like(x)
        inv = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
        L.call("eegclip_l2norm_forward", L.ptr(x), L.ptr(xn), L.ptr(inv), x.shape[0], x.shape[1], L.stream())
        return xn, inv

    def l2norm_bwd(self, xn, inv, dxn):
        dx = torch.empty_like(xn)
        L.call("eegclip_l2norm_backward", L.ptr(xn), L.ptr(inv), L.ptr(L.f32c(dxn)), L.ptr(dx), xn.shape[0], xn.shape[1], L.stream())
        return dx

    def lse(self, S_all, E_all, tau, b, row0, one_sided):
        Bg, D = E_all.shape
        dev = E_all.device
        out = torch.empty(3, b, dtype=torch.float32, device=dev)  # lse_row, lse_col, diag
        if one_sided:
            out[1].zero_()
        scratch = self._scratch(b, Bg, D, dev)
        L.call("eegclip_infonce_lse", L.ptr(S_all), L.ptr(E_all), L.ptr(tau), b, row0, Bg, D, L.ptr(out[0]), L.ptr(out[1]),
               L.ptr(out[2]), L.default_math(), int(one_sided), L.ptr(scratch), L.stream())
        return out

    def loss(self, vec_all, one_sided):
        Bg = vec_all.shape[1]
        loss = torch.empty((), dtype=torch.float32, device=vec_all.device)
        L.call("eegclip_infonce_loss", L.ptr(vec_all[0]), L.ptr(vec_all[1]), L.ptr(vec_all[2]), Bg, int(one_sided), L.ptr(loss),
               L.stream())
        return loss

    def backward(self, S_all, E_all, tau, vec_all, b, row0, dloss, one_sided):
        Bg, D = E_all.shape
        dev = E_all.device
        dS = None if one_sided else torch.empty(b, D, dtype=torch.float32, device=dev)
        dE = torch.empty(b, D, dtype=torch.float32, device=dev)
        dtau = torch.empty((), dtype=torch.float32, device=dev)
        scratch = self._scratch(b, Bg, D, dev)
        L.call("eegclip_infonce_backward", L.ptr(S_all), L.ptr(E_all), L.ptr(tau), L.ptr(vec_all[0]), L.ptr(vec_all[1]), b, row0,
               Bg, D, L.ptr(dloss), L.ptr(dS), L.ptr(dE), L.ptr(dtau), L.default_math(), int(one_sided), L.ptr(scratch), L.stream())
        return dS, dE, dtau


_CUDA_OPS = CudaHeadOps()


def _world(group):
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _all_gather_rows(x, group, world):
    """(b, ...) -> (world*b, ...) in rank order."""
    if world == 1:
        return x
    out = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=group)
    return out


class _InfoNCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, E_raw, S_raw, tau, group, ops):
        world, rank = _world(group)
        b = E_raw.shape[0]
        En, invE = ops.l2norm_fwd(E_raw)
        Sn, invS = ops.l2norm_fwd(S_raw)
        E_all = _all_gather_rows(En, group, world)
        S_all = _all_gather_rows(Sn, group, world)
        tau_c = tau.detach().to(En.dtype).reshape(1).contiguous()
        vec = ops.lse(S_all, E_all, tau_c, b, rank * b, False)              # (3, b)
        if world > 1:
            gathered = _all_gather_rows(vec.t().contiguous(), group, world)  # (world*b, 3)
            vec_all = gathered.t().contiguous()
        else:
            vec_all = vec
        loss = ops.loss(vec_all, False)
        ctx.saved = (En, invE, Sn, invS, E_all, S_all, tau_c, vec_all)
        ctx.meta = (b, rank * b, ops, tau.shape)
        E_all_out = E_all if world > 1 else En.detach()     # (a distinct tensor object: one output per object)
        ctx.mark_non_differentiable(En, E_all_out)
        return loss, En, E_all_out

    @staticmethod
    def backward(ctx, dloss, _dEn, _dEall):
        En, invE, Sn, invS, E_all, S_all, tau_c, vec_all = ctx.saved
        b, row0, ops, tau_shape = ctx.meta
        dl = dloss.detach().to(En.dtype).reshape(1).contiguous()
        dSn, dEn, dtau = ops.backward(S_all, E_all, tau_c, vec_all, b, row0, dl, False)
        dE = ops.l2norm_bwd(En, invE, dEn)
        dS = ops.l2norm_bwd(Sn, invS, dSn)
        return dE, dS, dtau.reshape(tau_shape), None, None


class _CERowsFn(torch.autograd.Function):
    """CE(normalize(X) . En^T * exp(tau), arange) with gradients for En and tau only (clip_model.py:919,934-937)."""

    @staticmethod
    def forward(ctx, X_raw, En, tau, ops):
        Xn, _ = ops.l2norm_fwd(X_raw)
        En = En.detach().contiguous()
        tau_c = tau.detach().to(En.dtype).reshape(1).contiguous()
        b = En.shape[0]
        vec = ops.lse(Xn, En, tau_c, b, 0, True)
        loss = ops.loss(vec, True)
        ctx.saved = (Xn, En, tau_c, vec)
        ctx.meta = (b, ops, tau.shape)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        Xn, En, tau_c, vec = ctx.saved
        b, ops, tau_shape = ctx.meta
        dl = dloss.detach().to(En.dtype).reshape(1).contiguous()
        _, dEn, dtau = ops.backward(Xn, En, tau_c, vec, b, 0, dl, True)
        return None, dEn, dtau.reshape(tau_shape), None


class _L2NormFn(torch.autograd.Function):
    """F.normalize(p=2, dim=1) with its backward (clip_model.py:913)."""

    @staticmethod
    def forward(ctx, x, ops):
        xn, inv = ops.l2norm_fwd(x)
        ctx.saved, ctx.ops = (xn, inv), ops
        return xn

    @staticmethod
    def backward(ctx, dxn):
        xn, inv = ctx.saved
        return ctx.ops.l2norm_bwd(xn, inv, dxn.contiguous()), None


def l2_normalize(x, ops=None):
    return _L2NormFn.apply(x, ops or _CUDA_OPS)


def infonce_loss(E_raw, S_raw, tau, group=None, return_normalized=False, ops=None):
    """Symmetric InfoNCE of clip_model.py:675-693 on raw (un-normalised) flattened embeddings (b,D).

    With ``group`` set and torch.distributed initialised the batch is the concatenation over ranks.
    Returns the loss (0-dim) and, if asked, this rank's normalised EEG embeddings (detached); ``return_normalized="all"`` adds the
    gathered normalised EEG embeddings of all ranks (what a sharded memory-bank update needs).
    """
    ops = ops or _CUDA_OPS
    if ops is _CUDA_OPS and not E_raw.is_cuda:
        raise L.EegclipError("infonce_loss: embeddings must be CUDA tensors (no CPU fallback on this path)")
    loss, En, E_all = _InfoNCEFn.apply(E_raw, S_raw, tau, group, ops)
    if return_normalized == "all":     # (loss, this rank's normalised EEG rows, the gathered normalised EEG rows of every rank)
        return loss, En, E_all
    return (loss, En) if return_normalized else loss


def ce_rows_loss(X_raw, En, tau, ops=None):
    return _CERowsFn.apply(X_raw, En, tau, ops or _CUDA_OPS)


# ---------------------------------------------------------------------------------------------------
# Data-parallel plumbing
# ---------------------------------------------------------------------------------------------------
def broadcast_parameters(module, group=None, src=0):
    """Make every rank start from rank ``src``'s weights and buffers."""
    world, _ = _world(group if group is not None else dist.group.WORLD if dist.is_initialized() else None)
    if world == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)


class BucketedGradReducer:
    """SUM-all-reduce of the optimizer's flat gradient arena in two buckets: the slice a tower backward has just finished writing is
    reduced asynchronously (NCCL stream) while the rest of the backward runs, the remainder after the backward.  Falls back to one
    collective over the whole arena whenever the written views do not form one exact contiguous slice."""

    def __init__(self, optimizer, group):
        self.opt, self.group = optimizer, group
        self.pending = []          # (flat index, lo, hi, work)

    def __enter__(self):
        L._SINK_LISTENERS.append(self._on_sinks)
        return self

    def __exit__(self, *exc):
        L._SINK_LISTENERS.remove(self._on_sinks)
        return False

    def _on_sinks(self, views):
        if not views or self.pending:
            return
        for fi, a in self.opt._arena.items():
            flat = a["flat_g"]
            base, esz = flat.data_ptr(), flat.element_size()
            los = [(v.data_ptr() - base) // esz for v in views]
            if min(los) < 0 or max(los) >= flat.numel():
                continue
            lo = min(los)
            hi = max(o + (v.numel() + 3) // 4 * 4 for o, v in zip(los, views))
            if hi > flat.numel() or sum((v.numel() + 3) // 4 * 4 for v in views) != hi - lo:
                return                                     # not one exact slice: leave it to the final collective
            work = dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.pending.append((fi, lo, hi, work))
            return

    def finish(self):
        """After backward: reduce what has not been reduced yet, then make the compute stream wait for the early bucket."""
        flats = self.opt.flat_grads()                      # (folds autograd-produced gradients into the arena first)
        done = {fi: (lo, hi) for fi, lo, hi, _ in self.pending}
        for fi, flat in enumerate(flats):
            if fi in done:
                lo, hi = done[fi]
                if lo > 0:
                    dist.all_reduce(flat[:lo], op=dist.ReduceOp.SUM, group=self.group)
                if hi < flat.numel():
                    dist.all_reduce(flat[hi:], op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        for _, _, _, work in self.pending:
            work.wait()
        self.pending = []


def allreduce_gradients(params, group=None, flat=None):
    """SUM-all-reduce the gradients (one collective when ``flat`` -- the optimizer's flat gradient arena -- is given)."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    if dist.get_world_size(group) == 1:
        return
    if flat is not None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    buf = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for g in grads:
        g.copy_(buf[o:o + g.numel()].view_as(g))
        o += g.numel()


def bind_to_gpu_numa_node(device_index):
    """Pin this process (and therefore the first-touch placement of the pinned staging buffers it allocates next) to the CPUs
    of the NUMA node the GPU hangs off.  With one process per GPU and 357 MB of fp32 speech features per rank and step, eight
    ranks that all stage from one socket's memory are bound by that socket's memory / UPI bandwidth, not by PCIe (round-1 SCALE:
    end-to-end efficiency 0.91 at 4 and 8 GPUs against 0.98 device-side).  Returns the node id, or None when the topology
    cannot be read (the call is then a no-op)."""
    import os
    try:
        import torch
        bdf = torch.cuda.get_device_properties(device_index).pci_bus_id if hasattr(torch.cuda.get_device_properties(device_index), "pci_bus_id") else None
        if bdf is None:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:                 # nvml prints a 32-bit domain, sysfs a 16-bit one
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
