"""Drop-in for the optimizer side of the reference's training step (optimizer half).

The reference trains with `torch.optim.AdamW` (`BaseExperiment._get_optim`, `src/experiment_types/_base_experiment.py:711-725`;
`src/configs/optimizer/adamw.yaml`: lr 7e-5, weight decay 1e-6, eps 1e-8, betas (0.9, 0.99); the Navier-Stokes experiment uses
lr 3e-4 / wd 1e-4) and lets Lightning clip the global gradient norm first (`gradient_clip_val: 1.0`,
`src/configs/trainer/default.yaml:10`, i.e. `torch.nn.utils.clip_grad_norm_`).  `AdamW` here keeps torch's constructor, `step`,
`zero_grad`, `state_dict` / `load_state_dict` (torch's layout, so optimizer states move between the two) and adds
`max_grad_norm`: parameters, gradients and both moments are views of four flat arenas and one `dyf_adamw_step` call does the
norm reduction, the clipping and the update (`csrc/optim_kernels.cu`) on the current stream, without a host round trip.

The backward kernels that would fill the gradient arena are not built yet; any producer of `p.grad` works (autograd
accumulates into the arena views in place).  No CPU / PyTorch fallback: parameters must be CUDA float32 tensors."""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Iterable, List, Optional, Sequence

import torch

from . import engine as E


class _Arena:
    """One parameter group laid out back to back (every tensor starts on a 16-byte boundary)."""

    def __init__(self, params: List[torch.nn.Parameter]):
        dev = params[0].device
        self.offsets, n = [], 0
        for p in params:
            self.offsets.append(n)
            n += (p.numel() + 3) // 4 * 4
        self.n = n
        self.params = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grads = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        nbytes = C.c_size_t()
        E._check(E.LIB.dyf_adamw_workspace_bytes(n, C.byref(nbytes)))
        self.workspace = torch.zeros(nbytes.value // 8, dtype=torch.float64, device=dev)
        self.step = 0
        self.step_tensor = torch.tensor(0.0)  # shared by every parameter's state["step"] (torch keeps one per parameter)
        self.grad_ptrs: List[int] = []        # data_ptr of every gradient view (checked per step without building views)

    def view(self, flat: torch.Tensor, i: int, like: torch.Tensor) -> torch.Tensor:
        o = self.offsets[i]
        return flat[o:o + like.numel()].view(like.shape)


class AdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 amsgrad: bool = False, max_grad_norm: Optional[float] = None, owners: Sequence[Any] = ()):
        if not 0.0 <= lr:  # torch/optim/adamw.py raises the same ValueErrors
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        if amsgrad:
            raise NotImplementedError("amsgrad is not built (the reference does not use it)")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      max_grad_norm=max_grad_norm))
        self._owners = list(owners)  # engine backbones whose packed weights must be rebuilt after a step (`mark_dirty`)
        self._arenas: List[_Arena] = []
        for group in self.param_groups:
            ps = group["params"]
            for p in ps:
                if not p.is_cuda:
                    raise E.EngineError("dyffusion_b200.optim has no CPU path: parameters must be CUDA tensors")
                if p.dtype != torch.float32:
                    raise ValueError("parameters must be float32 (the reference's bf16-mixed runs keep fp32 master weights)")
            arena = _Arena(ps)
            with torch.no_grad():
                for i, p in enumerate(ps):
                    v = arena.view(arena.params, i, p)
                    v.copy_(p.data)
                    p.data = v                                   # the module now reads / writes the arena
                    g = arena.view(arena.grads, i, p)
                    if p.grad is not None:
                        g.copy_(p.grad)
                    p.grad = g                                   # autograd accumulates in place into the arena
                    arena.grad_ptrs.append(g.data_ptr())
                    self.state[p] = {"step": arena.step_tensor, "exp_avg": arena.view(arena.exp_avg, i, p),
                                     "exp_avg_sq": arena.view(arena.exp_avg_sq, i, p)}
            self._arenas.append(arena)

    # ------------------------------------------------------------------ torch.optim.Optimizer surface
    def add_param_group(self, param_group) -> None:
        if getattr(self, "_arenas", None):  # torch calls this during construction; later additions cannot join an arena
            raise NotImplementedError("parameter groups are laid out in flat arenas at construction: build a new optimizer")
        super().add_param_group(param_group)

    def zero_grad(self, set_to_none: bool = False) -> None:
        """One memset per group; the gradient views stay attached (`set_to_none` would detach them and is ignored)."""
        for group, arena in zip(self.param_groups, self._arenas):
            arena.grads.zero_()
            for i, p in enumerate(group["params"]):
                if p.grad is None:
                    p.grad = arena.view(arena.grads, i, p)

    def _collect_grads(self, group, arena: _Arena) -> None:
        for i, (p, ptr) in enumerate(zip(group["params"], arena.grad_ptrs)):
            g = p.grad
            if g is not None and g.data_ptr() == ptr:
                continue          # the usual case: autograd accumulated in place into the arena view
            want = arena.view(arena.grads, i, p)
            if g is None:
                want.zero_()      # torch skips parameters without a gradient; a zero gradient still decays them --
            else:                 # documented difference (the reference's parameters all receive gradients)
                want.copy_(g)
            p.grad = want

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group, arena in zip(self.param_groups, self._arenas):
            self._collect_grads(group, arena)
            arena.step += 1
            b1, b2 = group["betas"]
            clip = group.get("max_grad_norm") or 0.0
            with torch.cuda.device(arena.params.device):
                E._check(E.LIB.dyf_adamw_step(arena.params.data_ptr(), arena.grads.data_ptr(), arena.exp_avg.data_ptr(),
                                              arena.exp_avg_sq.data_ptr(), arena.n, float(group["lr"]), float(b1), float(b2),
                                              float(group["eps"]), float(group["weight_decay"]), arena.step, float(clip),
                                              arena.workspace.data_ptr(), arena.workspace.numel() * 8, E._stream_ptr()))
            arena.step_tensor.fill_(float(arena.step))
        for o in self._owners:
            o.mark_dirty()
        return loss

    def sync_gradients(self, process_group=None) -> None:
        """DDP's gradient averaging as one all-reduce per arena (`dyffusion_b200.distributed.allreduce_mean_`); call it
        between `backward()` and `step()` -- the clip then sees the averaged gradients, as under Lightning's DDP."""
        from .distributed import allreduce_mean_
        for group, arena in zip(self.param_groups, self._arenas):
            self._collect_grads(group, arena)
            allreduce_mean_(arena.grads, process_group)

    def grad_norm(self, group: int = 0) -> torch.Tensor:
        """Global L2 norm of the gradients seen by the last clipped `step` (device scalar; Lightning logs `grad_norm`)."""
        return self._arenas[group].workspace[-1].sqrt()

    def load_state_dict(self, state_dict: Dict[str, Any]) -> None:
        """Accepts torch.optim.AdamW's state dict: moments are copied into the arenas, the views stay in place.
        torch replaces every param_group by the saved one (keeping only `params`), and a group saved by torch.optim.AdamW has
        no `max_grad_norm`: the clipping threshold of the live groups is kept in that case (it would silently vanish)."""
        clips = [g.get("max_grad_norm", self.defaults.get("max_grad_norm")) for g in self.param_groups]
        super().load_state_dict(state_dict)
        for g, clip in zip(self.param_groups, clips):
            if "max_grad_norm" not in g:
                g["max_grad_norm"] = clip
        for group, arena in zip(self.param_groups, self._arenas):
            steps = set()
            for i, p in enumerate(group["params"]):
                st = self.state.get(p, {})
                for name, flat in (("exp_avg", arena.exp_avg), ("exp_avg_sq", arena.exp_avg_sq)):
                    v = arena.view(flat, i, p)
                    if name in st and st[name].data_ptr() != v.data_ptr():
                        v.copy_(st[name])
                    st[name] = v
                steps.add(int(float(st.get("step", 0.0))))
                st["step"] = arena.step_tensor
                self.state[p] = st
            if len(steps) > 1:
                raise ValueError(f"parameters of one group carry different step counts: {sorted(steps)}")
            arena.step = steps.pop() if steps else 0
            arena.step_tensor.fill_(float(arena.step))
