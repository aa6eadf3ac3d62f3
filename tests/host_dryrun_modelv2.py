"""Host-logic dry run of the ModelV2 FCNet class (CPU, own process — it monkeypatches torch; run by tests/test_host.py).

`ddrl_b200.kernels.fcnet_forward / fcnet_backward` are replaced by the ORACLE's float64 forward and its autograd, `.cuda()`
becomes the identity, and the GPU tests of the class are executed unchanged: the default layout and the `vf_share_layers` /
`free_log_std` layouts served through `modelv2.fcnet_layout_map` (variable names and order, weight tying, constant columns,
gradient tying through autograd's index-select).  Checks no kernel; test infrastructure only."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle.ddrl_oracle as O  # noqa: E402

torch.cuda.is_available = lambda: True
torch.Tensor.cuda = lambda self, *a, **k: self

import ddrl_b200.kernels as K  # noqa: E402


def fwd(theta, obs, A, **k):
    lg, v = O.fcnet_forward(theta[0].double(), obs[0].double(), 2 * A)
    return {"logits": lg.float()[None], "value": v.float()[None]}


def bwd(theta, obs, dl, dv, A):
    with torch.enable_grad():
        t = theta[0].double().requires_grad_(True)
        lg, v = O.fcnet_forward(t, obs[0].double(), 2 * A)
        (g,) = torch.autograd.grad((lg * dl[0].double()).sum() + (v * dv[0].double()).sum(), t)
    return g.float()[None]


K.fcnet_forward, K.fcnet_backward = fwd, bwd
import tests.test_gpu_modelv2 as T  # noqa: E402
import tests.test_gpu_modelv2_layouts as Z  # noqa: E402

for flags in [(True, False), (False, True), (True, True)]:
    Z.test_fcnet_modelv2_optional_layouts(*flags)
    print("layouts", *flags, "ok")
T.test_fcnet_modelv2_forward_value_and_autograd()
print("default ok")
T.test_unsupported_model_configs_fail_loudly()
print("unsupported ok")
