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

    # ---- checkpoint format of torch.optim.Adam (utils.py:479-497 saves optimizer.state_dict()) ----
    def _names(self):
        by_id = {id(p): n for n, p in self.net.named_parameters()}
        return [by_id[id(p)] for p in self.param_groups[0]["params"]]

    def state_dict(self):
        """Same layout as torch.optim.Adam.state_dict(): per-parameter step / exp_avg / exp_avg_sq in the order the
        parameters were given, so a checkpoint written here resumes under the reference's Adam and vice versa."""
        st, state = self.net.store, {}
        for i, n in enumerate(self._names()):
            if self.t > 0 and st.offsets[n] < st.used:                 # parameters that never get a gradient have no state
                state[i] = {"step": torch.tensor(float(self.t)),
                            "exp_avg": st.view(n, self.m).contiguous().clone(),
                            "exp_avg_sq": st.view(n, self.v).contiguous().clone()}
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g.update(weight_decay=0, amsgrad=False, params=list(range(len(self.param_groups[0]["params"]))))
        return {"state": state, "param_groups": [g]}

    @torch.no_grad()
    def load_state_dict(self, sd):
        st, names = self.net.store, self._names()
        g = sd["param_groups"][0]
        if len(g["params"]) != len(names):
            raise ValueError(f"optimizer state has {len(g['params'])} parameters, the net has {len(names)}")
        if g.get("weight_decay", 0) or g.get("amsgrad", False):
            raise ValueError("FusedAdam implements Adam without weight decay / amsgrad (main_dist.py:50)")
        for k in ("lr", "betas", "eps"):
            self.param_groups[0][k] = g[k]
        self.m.zero_()
        self.v.zero_()
        steps = set()
        for i, s in sd["state"].items():
            n = names[int(i)]
            st.view(n, self.m).copy_(s["exp_avg"])
            st.view(n, self.v).copy_(s["exp_avg_sq"])
            steps.add(int(s["step"]))
        if len(steps) > 1:
            raise ValueError(f"per-parameter step counts differ ({sorted(steps)}): not a state this optimizer can resume")
        self.t = steps.pop() if steps else 0

    def zero_grad(self, set_to_none=True):
        # gradients live in the arena, which the engine clears at the start of every backward
        super().zero_grad(set_to_none=set_to_none)
