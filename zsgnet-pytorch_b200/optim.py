"""Adam as configured by the reference (main_dist.py:50: betas=(0.9, 0.99), utils.py:667-672), run as ONE
kernel over the flat parameter / gradient arenas instead of ~160 per-tensor updates."""
import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Optimizer-compatible (param_groups, state_dict, zero_grad) so that the reference's
    Learner and ReduceLROnPlateau scheduler (utils.py:674-691) drive it unchanged."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.99), eps=1e-8, net=None, reducer=None):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        if net is None:
            raise ValueError("FusedAdam needs net=<zsg_b200.mdl.ZSGNet> (it updates the net's parameter arena)")
        self.net, self.reducer = net, reducer
        st = net.store
        self.m = torch.zeros(st.used, device=st.device)
        self.v = torch.zeros(st.used, device=st.device)
        self.t = 0

    @torch.no_grad()
    def step(self, closure=None):
        st = self.net.store
        if self.reducer is not None:
            self.reducer.finish()
        g = self.param_groups[0]
        self.t += 1
        scale = self.reducer.grad_scale if self.reducer is not None else 1.0
        ops.adam(st.param_arena, st.grad_arena, self.m, self.v, st.used, float(g["lr"]), float(g["betas"][0]),
                 float(g["betas"][1]), float(g["eps"]), self.t, scale)

    def zero_grad(self, set_to_none=True):
        # gradients live in the arena, which the engine clears at the start of every backward
        super().zero_grad(set_to_none=set_to_none)
