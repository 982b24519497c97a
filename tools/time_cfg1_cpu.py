"""BASELINE.json config 1 ("2D pusher Jacobian-field forward (project/jacobian), 64x64, CPU-only PyTorch"): a CPU
sanity timing of the reference's own UnetJacobianField.forward (project/jacobian/models/jacobian_models/
unet_jacobian.py:52-66), imported UNMODIFIED from /root/reference in the build container (it cannot travel to the GPU
box).  Out of scope for kernels (dense conv net, no volumetric path); recorded for the
record in profiles/cfg1_cpu_r02.json."""
import json
import os
import sys
import time
import types

import torch

REF = "/root/reference/project"
sys.path.insert(0, REF)
# the package's __init__ chain pulls in lightning / hydra pieces that are absent here: import the two modules directly
for name in ("jacobian", "jacobian.models", "jacobian.models.jacobian_models", "jacobian.model_components"):
    m = types.ModuleType(name)
    m.__path__ = [os.path.join(REF, *name.split("."))]
    sys.modules[name] = m
oc = types.ModuleType("omegaconf")   # type annotation only (base_jacobian.py:7)
oc.DictConfig = dict
sys.modules.setdefault("omegaconf", oc)
from jacobian.models.jacobian_models.unet_jacobian import UnetJacobianField, UnetJacobianFieldCfg  # noqa: E402


def main():
    torch.manual_seed(0)
    cfg = UnetJacobianFieldCfg(name="unet", command_dim=2, spatial_dim=2) if "command_dim" in UnetJacobianFieldCfg.__dataclass_fields__ \
        else UnetJacobianFieldCfg()
    net = UnetJacobianField(cfg).eval()
    img, cmd = torch.rand(1, 3, 64, 64), torch.rand(1, 2)
    with torch.no_grad():
        for _ in range(5):
            out = net(img, cmd)
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            out = net(img, cmd)
        dt = (time.perf_counter() - t0) / n
    res = {"config": "cfg1: 2D pusher Jacobian-field forward, 64x64, CPU-only PyTorch", "ms_per_forward": dt * 1e3,
           "forwards_per_s": 1.0 / dt, "threads": torch.get_num_threads(), "params": sum(p.numel() for p in net.parameters()),
           "jacobian_shape": list(out.jacobian.shape), "flow_shape": list(out.flow.shape), "where": "build container (no GPU)",
           "source": "unmodified /root/reference/project/jacobian/models/jacobian_models/unet_jacobian.py"}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
