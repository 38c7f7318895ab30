"""Host-side handle of the tcgen05 evaluator network (csrc/evalnet.cu, C ABI sprl_evalnet_*).

The reference's worker loads whatever TorchScript file the controller traced and runs it with
LibTorch (cpp/src/networks/GridNetwork.hpp:37-51,99).  For 8x8 boards this handle takes the
same module's parameters (by the names of src/networks/grid_networks.py:30-80) and runs the
forward pass in libsprl_b200.so instead; torch is used only to read the state_dict."""
import ctypes as C

import numpy as np

from . import capi


def _f32(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if hasattr(t, "detach") else t, dtype=np.float32)


def network_params(state_dict, rows=8, cols=8):
    """Builds a capi.NetworkParams (host pointers) from a BasicGridNetwork state_dict.
    Returns (params, keepalive) -- keep `keepalive` referenced while `params` is in use."""
    sd = {k: _f32(v) for k, v in state_dict.items() if not k.endswith("num_batches_tracked")}
    keep = [sd]

    def conv_bn(conv, bn):
        cb = capi.ConvBnParams()
        for field, key in (("weight", conv + ".weight"), ("bias", conv + ".bias"), ("bn_weight", bn + ".weight"),
                           ("bn_bias", bn + ".bias"), ("bn_mean", bn + ".running_mean"), ("bn_var", bn + ".running_var")):
            setattr(cb, field, sd[key].ctypes.data)
        return cb

    blocks = 0
    while f"residual_blocks.{blocks}.conv1.weight" in sd:
        blocks += 1
    p = capi.NetworkParams()
    p.rows, p.cols = rows, cols
    p.channels, p.in_planes = sd["conv.weight"].shape[0], sd["conv.weight"].shape[1]
    p.blocks = blocks
    p.actions = sd["policy_fc.weight"].shape[0]
    p.policy_channels = sd["policy_conv.weight"].shape[0]
    p.value_channels = sd["value_conv.weight"].shape[0]
    p.value_hidden = sd["value_fc1.weight"].shape[0]
    p.bn_eps = 1e-5
    p.stem = conv_bn("conv", "bn")
    tower = (capi.ConvBnParams * max(1, 2 * blocks))()
    for b in range(blocks):
        tower[2 * b] = conv_bn(f"residual_blocks.{b}.conv1", f"residual_blocks.{b}.bn1")
        tower[2 * b + 1] = conv_bn(f"residual_blocks.{b}.conv2", f"residual_blocks.{b}.bn2")
    keep.append(tower)
    p.tower = C.cast(tower, C.POINTER(capi.ConvBnParams))
    for field in ("policy_conv", "policy_fc", "value_conv", "value_fc1", "value_fc2"):
        setattr(p, field + "_w", sd[field + ".weight"].ctypes.data)
        setattr(p, field + "_b", sd[field + ".bias"].ctypes.data)
    return p, keep


class EvalNet:
    """sprl_evalnet handle: forward(d_in, batch, d_logits, d_value, stream) on device pointers."""

    def __init__(self, module_or_state_dict, device=0, rows=8, cols=8):
        self.lib = capi.load()
        sd = module_or_state_dict.state_dict() if hasattr(module_or_state_dict, "state_dict") else module_or_state_dict
        params, keep = network_params(sd, rows, cols)
        self.handle = C.c_void_p()
        capi.check(self.lib.sprl_evalnet_create(device, C.byref(params), C.byref(self.handle)))
        self.device, self.rows, self.cols = device, rows, cols
        self.actions, self.in_planes = params.actions, params.in_planes
        self.weight_bytes = sum(v.nbytes for v in keep[0].values())
        del keep

    def update(self, module_or_state_dict):
        sd = module_or_state_dict.state_dict() if hasattr(module_or_state_dict, "state_dict") else module_or_state_dict
        params, keep = network_params(sd, self.rows, self.cols)
        capi.check(self.lib.sprl_evalnet_update(self.handle, C.byref(params)))
        del keep

    def forward_ptr(self, d_in, batch, d_logits, d_value, stream=0, d_rows=None):
        """d_rows: device address of a uint32 holding the number of rows to evaluate (<= batch), read by the
        kernel at run time -- the engine's compact leaf batch (sprl_eval_rows); None = all `batch` rows."""
        if d_rows:
            capi.check(self.lib.sprl_evalnet_forward_counted(self.handle, C.c_void_p(d_in), C.c_void_p(d_rows), batch,
                                                              C.c_void_p(d_logits), C.c_void_p(d_value), C.c_void_p(stream)))
        else:
            capi.check(self.lib.sprl_evalnet_forward(self.handle, C.c_void_p(d_in), batch, C.c_void_p(d_logits),
                                                      C.c_void_p(d_value), C.c_void_p(stream)))

    def __call__(self, x):
        """Torch convenience: x [B, planes, 8, 8] fp32 on this GPU -> (logits [B, A], value [B, 1])."""
        import torch
        x = x.contiguous()
        logits = torch.empty((x.shape[0], self.actions), dtype=torch.float32, device=x.device)
        value = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
        self.forward_ptr(x.data_ptr(), x.shape[0], logits.data_ptr(), value.data_ptr(),
                         torch.cuda.current_stream(x.device).cuda_stream)
        return logits, value.reshape(-1, 1)

    def set_precision(self, precision):
        """capi.EVALNET_PRECISION_FP32_SPLIT (default: fp32-level accuracy, three fp16 MMAs per product) or
        capi.EVALNET_PRECISION_FP16 (opt-in: one MMA per product, ~1e-3 on logits; not the reference's precision)."""
        capi.check(self.lib.sprl_evalnet_set_precision(self.handle, precision))

    def set_path(self, path):
        """capi.EVALNET_PATH_AUTO / _STREAMING / _RESIDENT: which conv-tower kernel runs (same arithmetic, same order)."""
        capi.check(self.lib.sprl_evalnet_set_path(self.handle, path))

    @property
    def phases(self):
        """Launches of the resident-weight kernel per forward; 0 when the streaming kernel is in use."""
        return self.lib.sprl_evalnet_phases(self.handle)

    def info(self):
        up, st, sm = C.c_int64(), C.c_int32(), C.c_int32()
        capi.check(self.lib.sprl_evalnet_info(self.handle, C.byref(up), C.byref(st), C.byref(sm)))
        return dict(upload_bytes=up.value, ring_stages=st.value, smem_bytes=sm.value)

    def status(self):
        n = C.c_uint64()
        capi.check(self.lib.sprl_evalnet_status(self.handle, C.byref(n)))
        return n.value

    def close(self):
        if self.handle:
            self.lib.sprl_evalnet_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
