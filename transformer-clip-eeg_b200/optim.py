"""AdamW on one multi-tensor kernel (eegclip_adamw_step), torch.optim.AdamW-compatible surface.

Replaces ``torch.optim.AdamW(model.parameters(), betas, amsgrad, weight_decay, lr)`` of
train_clip_final.py:409-413,492.  Gradients live in one persistent flat arena (``flat_grad``): parameters'
``.grad`` are views into it, ``zero_grad`` is one memset, the data-parallel all-reduce is one NCCL call over
the arena, and the device-side tensor table is uploaded once.
"""
import torch

from . import _lib as L


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, amsgrad=False):
        if amsgrad:
            raise L.EegclipError("amsgrad is not implemented on the B200 AdamW kernel (reference default is off)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._arena = {}

    def _group_arena(self, gi, group):
        """Flat grad / exp_avg / exp_avg_sq arenas and the device tensor table for one param group."""
        ps = [p for p in group["params"] if p.requires_grad]
        key = tuple((p.data_ptr(), p.numel()) for p in ps)
        a = self._arena.get(gi)
        if a is not None and a["key"] == key:
            return a
        dev = ps[0].device
        if dev.type != "cuda":
            raise L.EegclipError("AdamW: parameters must live on a CUDA device (no CPU fallback on this path)")
        offs, o = [], 0
        for p in ps:
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        flat_g = torch.zeros(o, dtype=torch.float32, device=dev)
        flat_m = torch.zeros(o, dtype=torch.float32, device=dev)
        flat_v = torch.zeros(o, dtype=torch.float32, device=dev)
        rows = []
        for p, off in zip(ps, offs):
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise L.EegclipError("AdamW: parameters must be contiguous fp32")
            view = flat_g[off:off + p.numel()].view(p.shape)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            rows.append([p.data_ptr(), view.data_ptr(), flat_m.data_ptr() + 4 * off, flat_v.data_ptr() + 4 * off, p.numel()])
        table = torch.tensor(rows, dtype=torch.int64).to(dev)
        a = dict(key=key, ps=ps, flat_g=flat_g, flat_m=flat_m, flat_v=flat_v, table=table, max_numel=max(p.numel() for p in ps),
                 step=0, written=set(p.data_ptr() for p in ps))   # nothing may be written directly before the first zero_grad
        L.register_grad_sinks(ps, [p.grad for p in ps], a)
        self._arena[gi] = a
        return a

    def flat_grads(self):
        return [self._group_arena(gi, g)["flat_g"] for gi, g in enumerate(self.param_groups)]

    def zero_grad(self, set_to_none=True):
        # gradients stay views of the arena (stable pointers); one memset per group
        for gi, g in enumerate(self.param_groups):
            a = self._group_arena(gi, g)
            a["flat_g"].zero_()
            a["written"].clear()                       # the package's backward kernels may now write each view once, directly

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for gi, g in enumerate(self.param_groups):
            a = self._group_arena(gi, g)
            # a gradient produced outside the arena (autograd replaced .grad): fold it back in
            for p in a["ps"]:
                if p.grad is None:
                    raise L.EegclipError("AdamW.step: a parameter has no gradient (reference loop always backprops all)")
            rebuilt = False
            for p in a["ps"]:
                st = p.grad.untyped_storage().data_ptr()
                if st != a["flat_g"].untyped_storage().data_ptr():
                    rebuilt = True
            if rebuilt:
                off = 0
                for p in a["ps"]:
                    view = a["flat_g"][off:off + p.numel()].view(p.shape)
                    if p.grad.data_ptr() != view.data_ptr():
                        view.copy_(p.grad)
                        p.grad = view
                    off += (p.numel() + 3) // 4 * 4
            a["step"] += 1
            b1, b2 = g["betas"]
            L.call("eegclip_adamw_step", L.ptr(a["table"]), len(a["ps"]), a["max_numel"], float(g["lr"]), float(b1), float(b2),
                   float(g["eps"]), float(g["weight_decay"]), a["step"], L.stream())
        return loss


Adam = None  # the reference's 'adam' branch (train_clip_final.py:403-407) is not on the default path
