"""numpy restatement of torch.optim.Adam as the reference configures it (TEST INFRASTRUCTURE ONLY):
optim.Adam(model.parameters(), lr=5e-5, weight_decay=1e-4), train_pad_20.py:54 - coupled L2 decay,
bias-corrected first/second moments, eps added outside the square root, parameters whose gradient is
None are skipped entirely (no decay, no step count)."""
import numpy as np


class AdamOracle:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.p = params                      # dict name -> ndarray, updated in place
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, betas[0], betas[1], eps, weight_decay
        self.state = {}

    def step(self, grads):
        for k, g in grads.items():
            if g is None:
                continue
            st = self.state.setdefault(k, dict(step=0, m=np.zeros_like(self.p[k]), v=np.zeros_like(self.p[k])))
            st["step"] += 1
            g = g + self.wd * self.p[k]
            st["m"] = self.b1 * st["m"] + (1 - self.b1) * g
            st["v"] = self.b2 * st["v"] + (1 - self.b2) * g * g
            bc1, bc2 = 1 - self.b1 ** st["step"], 1 - self.b2 ** st["step"]
            self.p[k] -= (self.lr / bc1) * st["m"] / (np.sqrt(st["v"]) / np.sqrt(bc2) + self.eps)
