"""What compute-sanitizer runs (tools/gpu_job.sh sanitize): smoke() -- marching, the fused training step, the inference
waves with fused compositing and ray compaction -- plus one small forward + backward of the wide feature head, so the TMA
GEMM kernels (forward, masked dgrad, window outputs, weight gradients with a partial last chunk) are covered too."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g

g.smoke()
from autolabel_b200 import tcnn

torch.manual_seed(1)
net = tcnn.Network(15, 512, {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "None",
                             "n_neurons": 512, "n_hidden_layers": 2}).cuda()
x = torch.randn(300, 15, device='cuda').requires_grad_(True)
y = net(x)
y.backward(torch.randn_like(y) * 1e-4)
torch.cuda.synchronize()
assert torch.isfinite(y).all().item() and torch.isfinite(net.params.grad).all().item() and torch.isfinite(x.grad).all().item()
print("wide head ok")
