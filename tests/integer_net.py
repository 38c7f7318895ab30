"""A network whose outputs are exact in fp32 on any device (test infrastructure, shared by the GPU parity tests, the
oracle tests and tests/golden/make_golden.py): integer weights on 0/1 planes, so every partial sum is a small integer
and the summation order cannot matter.  Run by torch on the GPU for the engine (the path a traced module takes: planes
written by the search kernel, logits / value read back in place), by numpy for the oracle's callback evaluator, and as a
TorchScript file by the verbatim reference's LibTorch GridNetwork (oracle/_ref/ref_trace_torch)."""
import numpy as np


class IntegerNet:
    def __init__(self, planes, cells, actions, seed=0):
        rs = np.random.RandomState(seed)
        self.wp = rs.randint(-3, 4, size=(planes * cells, actions)).astype(np.float32)
        self.wv = rs.randint(-2, 3, size=(planes * cells,)).astype(np.float32)
        self.calls = 0

    def numpy(self, x):
        flat = np.asarray(x, np.float32).reshape(x.shape[0], -1)
        return (flat @ self.wp) * np.float32(0.125), np.clip((flat @ self.wv) * np.float32(1.0 / 64), -1, 1)

    def torch_fn(self, dev):
        import torch
        torch.backends.cuda.matmul.allow_tf32 = False
        wp, wv = torch.from_numpy(self.wp).to(dev), torch.from_numpy(self.wv).to(dev)

        def fn(x):
            self.calls += 1
            flat = x.reshape(x.shape[0], -1)
            return (flat @ wp) * 0.125, torch.clamp((flat @ wv) * (1.0 / 64), -1, 1)
        return fn

    def save_torchscript(self, path, planes, rows, cols):
        """The same function as a traced module with the (logits [B, A], value [B, 1]) outputs GridNetwork.hpp:99-102 reads."""
        import torch

        wp, wv = torch.from_numpy(self.wp), torch.from_numpy(self.wv)

        class M(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.register_buffer("wp", wp)
                self.register_buffer("wv", wv)

            def forward(self, x):
                flat = x.reshape(x.shape[0], -1)
                return (flat @ self.wp) * 0.125, torch.clamp((flat @ self.wv) * (1.0 / 64), -1, 1).reshape(-1, 1)

        with torch.no_grad():
            torch.jit.trace(M().eval(), torch.zeros(2, planes, rows, cols)).save(path)
