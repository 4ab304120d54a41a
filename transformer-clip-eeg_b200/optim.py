"""AdamW / Adam on one multi-tensor kernel (eegclip_adamw_step), torch.optim-compatible surface.

Replaces ``torch.optim.AdamW(model.parameters(), betas, amsgrad, weight_decay, lr)`` and ``torch.optim.Adam(...)`` of
train_clip_final.py:403-413,492.  Gradients live in one persistent flat arena (``flat_grad``): the package's backward
kernels write parameter gradients straight into it ("gradient sinks"), ``zero_grad`` is one memset, the data-parallel
all-reduce is one NCCL call over the arena, and the device-side tensor table is uploaded once.

torch semantics kept:
  * ``zero_grad()`` sets ``.grad`` to None (set_to_none); a parameter that receives no gradient in a step keeps ``grad is
    None`` and is skipped by ``step()`` -- no weight decay, no moment update (e.g. ``temperature_eeg`` while the warm-up
    epochs back-propagate ``loss_ce`` only);
  * ``state_dict()`` / ``load_state_dict()`` carry ``step``, ``exp_avg``, ``exp_avg_sq`` (and ``max_exp_avg_sq``) per
    parameter in torch's layout; the tensors are views of the flat moment arenas;
  * moving the parameters after construction (``model.to(...)``) migrates the moments instead of resetting them.
Gradient sinks are only valid with ``loss.backward()`` followed by this optimizer: ``torch.autograd.grad`` and DDP hooks
see no parameter gradients for sunk parameters (the backward returns None for them).
"""
import torch

from . import _lib as L


class AdamW(torch.optim.Optimizer):
    _coupled_decay = False        # AdamW: p *= 1 - lr*wd;  Adam: g += wd*p

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, amsgrad=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=bool(amsgrad)))
        self._arena = {}

    # ---- arenas ------------------------------------------------------------------------------------
    def _group_arena(self, gi, group):
        """Flat grad / exp_avg / exp_avg_sq arenas and the per-parameter views for one param group."""
        ps = [p for p in group["params"] if p.requires_grad]
        key = tuple((p.data_ptr(), p.numel()) for p in ps)
        a = self._arena.get(gi)
        if a is not None and a["key"] == key:
            return a
        dev = ps[0].device
        if dev.type != "cuda":
            raise L.EegclipError("AdamW: parameters must live on a CUDA device (no CPU fallback on this path)")
        if a is not None and [n for _, n in a["key"]] != [n for _, n in key]:
            raise L.EegclipError("AdamW: the parameter list changed shape after the optimizer state was created")
        offs, o = [], 0
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise L.EegclipError("AdamW: parameters must be contiguous fp32")
            offs.append(o)
            o += (p.numel() + 3) // 4 * 4
        n_arena = 4 if group["amsgrad"] else 3
        flats = [torch.zeros(o, dtype=torch.float32, device=dev) for _ in range(n_arena)]
        step = 0
        if a is not None:                                  # same tensors at new addresses (model.to / .float()): keep the moments
            for new, old in zip(flats[1:], a["flats"][1:]):
                new.copy_(old)
            step = a["step"]
        views = [[f[off:off + p.numel()].view(p.shape) for p, off in zip(ps, offs)] for f in flats]
        for p, v in zip(ps, views[0]):
            if p.grad is not None:                         # a gradient that already exists moves into the arena
                v.copy_(p.grad)
                p.grad = v
        a = dict(key=key, ps=ps, offs=offs, flats=flats, flat_g=flats[0], views=views, step=step, table=None, active=None,
                 max_numel=max(p.numel() for p in ps),
                 written=set(p.data_ptr() for p in ps))    # nothing may be written directly before the first zero_grad
        L.register_grad_sinks(ps, views[0], a)
        self._arena[gi] = a
        self._publish_state(a, group)
        return a

    def _publish_state(self, a, group):
        """torch layout of the per-parameter state (views of the arenas; ``step`` as a 0-dim float tensor)."""
        for i, p in enumerate(a["ps"]):
            st = {"step": torch.tensor(float(a["step"])), "exp_avg": a["views"][1][i], "exp_avg_sq": a["views"][2][i]}
            if group["amsgrad"]:
                st["max_exp_avg_sq"] = a["views"][3][i]
            self.state[p] = st

    def _fold_foreign_grads(self, a):
        """Gradients that autograd produced outside the arena (e.g. the log-scales, whose gradient is returned by an autograd
        Function instead of being sunk) are copied into their arena views; returns the per-parameter "has a gradient" flags."""
        active = []
        for p, v in zip(a["ps"], a["views"][0]):
            if p.grad is None:                             # torch skips parameters without a gradient
                active.append(False)
                continue
            if p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
                p.grad = v
            active.append(True)
        return tuple(active)

    def flat_grads(self):
        """The flat gradient arenas (one per group) with EVERY existing gradient inside: what the data-parallel all-reduce sums."""
        out = []
        for gi, g in enumerate(self.param_groups):
            a = self._group_arena(gi, g)
            self._fold_foreign_grads(a)
            out.append(a["flat_g"])
        return out

    def state_dict(self):
        for gi, g in enumerate(self.param_groups):
            self._publish_state(self._group_arena(gi, g), g)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)                # fills self.state[p] with loaded copies
        loaded = {p: dict(st) for p, st in self.state.items()}   # (arena creation below republishes self.state)
        for gi, g in enumerate(self.param_groups):
            a = self._group_arena(gi, g)
            names = ["exp_avg", "exp_avg_sq"] + (["max_exp_avg_sq"] if g["amsgrad"] else [])
            for i, p in enumerate(a["ps"]):
                st = loaded.get(p)
                if not st:
                    continue
                for k, name in enumerate(names):
                    if name in st:
                        a["views"][1 + k][i].copy_(st[name])
                a["step"] = int(float(st.get("step", a["step"])))
            self._publish_state(a, g)

    # ---- torch.optim surface -----------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        # one memset per group; the package's backward kernels may now write each arena view once, directly
        for gi, g in enumerate(self.param_groups):
            a = self._group_arena(gi, g)
            a["flat_g"].zero_()
            a["written"].clear()
            for p, v in zip(a["ps"], a["views"][0]):
                p.grad = None if set_to_none else v

    def _table(self, a, active, amsgrad):
        if a["table"] is None or a["active"] != active:
            rows = []
            for i, p in enumerate(a["ps"]):
                if not active[i]:
                    continue
                v = a["views"]
                rows.append([p.data_ptr(), v[0][i].data_ptr(), v[1][i].data_ptr(), v[2][i].data_ptr(), p.numel(),
                             v[3][i].data_ptr() if amsgrad else 0])
            a["table"] = torch.tensor(rows, dtype=torch.int64).to(a["flat_g"].device) if rows else None
            a["active"] = active
            a["n_active"] = len(rows)
        return a["table"]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for gi, g in enumerate(self.param_groups):
            a = self._group_arena(gi, g)
            table = self._table(a, self._fold_foreign_grads(a), g["amsgrad"])
            a["step"] += 1
            if table is None:
                continue
            b1, b2 = g["betas"]
            L.call("eegclip_adamw_step", L.ptr(table), a["n_active"], a["max_numel"], float(g["lr"]), float(b1), float(b2),
                   float(g["eps"]), float(g["weight_decay"]), int(self._coupled_decay), a["step"], L.stream())
        return loss


class Adam(AdamW):
    """torch.optim.Adam (train_clip_final.py:403-407; also the regression head's optimizer, helpers:626): L2 decay is added
    to the gradient (default 0)."""
    _coupled_decay = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
