"""Wide feature head (16 -> 512 -> 512 -> 512, the C5 shape) forward + backward at a C5-sized sample count: CUDA-event
time per pass and the achieved tensor throughput.  Run under `ncu --metrics gpu__time_duration.sum -k regex:k_gemm`
for the per-GEMM split.   python tools/bench_wide.py [rows]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autolabel_b200 import tcnn

n = int(sys.argv[1]) if len(sys.argv) > 1 else 344192
net = tcnn.Network(15, 512, {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "None",
                             "n_neurons": 512, "n_hidden_layers": 2}).cuda()
x = torch.randn(n, 15, device='cuda').requires_grad_(True)
gy = torch.randn(n, 512, device='cuda') * 1e-4
flops_f = 2.0 * n * (16 * 512 + 2 * 512 * 512)


def step():
    y = net(x)
    y.backward(gy)


for _ in range(3):
    step()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
it = 10
tf = tb = 0.0
for _ in range(it):
    e[0].record()
    y = net(x)
    e[1].record()
    y.backward(gy)
    e[2].record()
    torch.cuda.synchronize()
    tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
tf /= it; tb /= it
print(f"rows {n}: forward {tf:.3f} ms ({flops_f / tf / 1e9:.0f} TFLOP/s)  backward {tb:.3f} ms ({2 * flops_f / tb / 1e9:.0f} TFLOP/s)")

# ---- the last forward layer as the field uses it: fp32 window into vals [n, 1 + 3 + C + F] at column 4 + C, fp16 ReLU
# copy into semo_in [n, F + 16]; variants isolate the cost of each output
from autolabel_b200 import _lib
from autolabel_b200._lib import call, ptr, stream_ptr
p = net.params.detach()
xh = torch.zeros(n, 16, dtype=torch.float16, device='cuda')
xh[:, :15] = x.detach().half()
ws = torch.empty(_lib.lib.al_mlp_wide_workspace(16, 512, 512, 2, n, 0), dtype=torch.uint8, device='cuda')
st = stream_ptr(torch.device('cuda'))


def fwd(o0, ld0, c0, h0, ldh):
    call("al_mlp_wide_forward", 16, 512, 512, 2, ptr(p), ptr(xh), 16, n, None,
         ptr(o0) if o0 is not None else None, ld0, c0, 0, 512 if o0 is not None else 0, 0,
         None, 0, 0, 0, 0, 0,
         ptr(h0) if h0 is not None else None, ldh, 0, 0, 512 if h0 is not None else 0, 1, ptr(ws), st)


dense = torch.empty(n, 512, device='cuda')
vals = torch.empty(n, 518, device='cuda')
vals8 = torch.empty(n, 520, device='cuda')
semo_in = torch.empty(n, 528, dtype=torch.float16, device='cuda')
for name, args_ in [("dense fp32", (dense, 512, 0, None, 0)), ("vals[:, 6:518] (ld 518)", (vals, 518, 6, None, 0)),
                    ("vals[:, 8:520] (ld 520)", (vals8, 520, 8, None, 0)), ("fp16 relu copy only", (None, 0, 0, semo_in, 528)),
                    ("vals ld 518 + fp16 copy (the field's call)", (vals, 518, 6, semo_in, 528))]:
    for _ in range(2):
        fwd(*args_)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fwd(*args_)
    e1.record()
    torch.cuda.synchronize()
    print(f"  forward, last layer -> {name}: {e0.elapsed_time(e1) / 5:.3f} ms (3 layers)")
