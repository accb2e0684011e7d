"""Debug aid: tcgen05 MLP forward / backward on one shape against the mma.sync back end, printing errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autolabel_b200 import _lib, tcnn

def run(backend, n_in, n_out, hidden, nh, n, do_bwd):
    _lib.lib.al_set_mlp_backend(backend)
    net = tcnn.Network(n_in, n_out, {"n_neurons": hidden, "n_hidden_layers": nh}).cuda()
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, n_in, generator=g).cuda().requires_grad_(True)
    y = net(x)
    torch.cuda.synchronize()
    res = [y.detach().clone()]
    if do_bwd:
        gy = torch.randn(n, n_out, generator=g).cuda() * 1e-3
        y.backward(gy)
        torch.cuda.synchronize()
        res += [x.grad.clone(), net.params.grad.clone()]
    return res

for shape in [(44, 16, 128, 2), (31, 3, 128, 2), (15, 64, 64, 2), (79, 2, 64, 1)]:
    for n in [128, 1000, 100000]:
        for do_bwd in [False, True]:
            ref = run(0, *shape, n, do_bwd)
            print(shape, n, "bwd" if do_bwd else "fwd", "launching tc ...", flush=True)
            out = run(1, *shape, n, do_bwd)
            for name, a, b in zip(["y", "dx", "dW"], out, ref):
                err = (a - b).abs().max().item(); sc = b.abs().max().item()
                print(f"   {name}: max abs diff {err:.3e} (scale {sc:.3e})  nan={torch.isnan(a).any().item()}", flush=True)
