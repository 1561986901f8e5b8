"""Keras-2.7-compatible Adam on flat device vectors (replaces keras.optimizers.Adam used at src/ExecutionRun.py:226
and optimizer.apply_gradients at src/NeRF.py:164,167).  Defaults are Keras': beta_1 0.9, beta_2 0.999, epsilon 1e-7.

The reference wraps Adam in a dynamic LossScaleOptimizer because it computes in float16; this path accumulates in
fp32 with fp32 master weights, so no loss scaling exists.
"""
import torch

from ._lib import call, ptr


class Adam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.learning_rate = float(learning_rate)
        self.beta_1, self.beta_2, self.epsilon = float(beta_1), float(beta_2), float(epsilon)
        self.iterations = 0
        self._m = None
        self._v = None

    def apply_flat(self, params_list, flat_grads):
        """params_list: flat fp32 device tensors; flat_grads: their gradients concatenated in the same order."""
        total = sum(p.numel() for p in params_list)
        if self._m is None:
            self._m = torch.zeros(total, dtype=torch.float32, device=flat_grads.device)
            self._v = torch.zeros(total, dtype=torch.float32, device=flat_grads.device)
        self.iterations += 1
        off = 0
        for p in params_list:
            n = p.numel()
            call("nerf_adam_step", ptr(p), ptr(flat_grads[off:off + n]), ptr(self._m[off:off + n]),
                 ptr(self._v[off:off + n]), n, self.learning_rate, self.beta_1, self.beta_2, self.epsilon,
                 self.iterations)
            off += n

    def _state(self, total, device):
        if self._m is None:
            self._m = torch.zeros(total, dtype=torch.float32, device=device)
            self._v = torch.zeros(total, dtype=torch.float32, device=device)

    def apply_one(self, param, grad, offset, total, t):
        """Update ONE of the flat parameter tensors (its moments live at ``offset`` of the shared state) as step ``t``;
        the caller advances ``iterations`` once per step.  Lets the fine network's update run on a side stream as soon
        as its gradients are final."""
        self._state(total, grad.device)
        n = param.numel()
        call("nerf_adam_step", ptr(param), ptr(grad), ptr(self._m[offset:offset + n]), ptr(self._v[offset:offset + n]), n,
             self.learning_rate, self.beta_1, self.beta_2, self.epsilon, int(t))

    def state_dict(self):
        return {"iterations": self.iterations, "m": self._m, "v": self._v, "learning_rate": self.learning_rate}
