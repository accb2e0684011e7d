"""Wide MLP heads (tiled tcgen05 GEMM path, csrc/gemm_tc.cu) against the fp32 oracle and the stated fp16 arithmetic:
the 512-d LSeg feature head (16 -> 512 -> 512 -> 512), its semantic head (528 -> 64 -> 2), the ScanNet label-set head
(80 -> 64 -> 606) and a 256-wide variant; forward, input gradient and weight gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # (n_in, n_out, hidden, n_hidden)
    (15, 512, 512, 2), (527, 2, 64, 1), (79, 606, 64, 1), (15, 256, 256, 2), (60, 16, 256, 2),
]


@pytest.mark.parametrize("n_in,n_out,hidden,n_hidden", SHAPES)
@pytest.mark.parametrize("n", [1, 200, 5000 + 37])
def test_wide_mlp_forward_backward(n_in, n_out, hidden, n_hidden, n):
    from autolabel_b200 import _lib, tcnn
    from oracle import field_oracle as fo
    from tests.helpers import record, rel_l2, rel_max
    net = tcnn.Network(n_in, n_out, {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "None",
                                     "n_neurons": hidden, "n_hidden_layers": n_hidden}).cuda()
    assert _lib.lib.al_mlp_num_params(net.in_pad, hidden, net.out_pad, n_hidden) < 0, "must exercise the wide path"
    g = torch.Generator().manual_seed(n + n_in)
    x = torch.randn(n, n_in, generator=g).cuda()
    x1 = x.clone().requires_grad_(True)
    y = net(x1)
    p2 = net.params.detach().clone().requires_grad_(True)
    x2 = x.clone().requires_grad_(True)
    oy = fo.mlp(x2, p2, net.in_pad, hidden, net.out_pad, n_hidden)[:, :n_out]
    assert y.shape == oy.shape
    e_y = (y - oy).abs().max().item()
    assert e_y < 2e-3 * max(1.0, oy.abs().max().item()), e_y     # K up to 528 products of fp16-rounded operands
    gy = torch.randn(n, n_out, generator=g).cuda() * 1e-4
    y.backward(gy)
    oy.backward(gy)
    sc = fo.grad_scale_for(gy.abs().max().item())
    ym, dxm, dWm = fo.mlp_fp16_model(x, net.params.detach(), net.in_pad, hidden, net.out_pad, n_hidden, dout=gy, scale=sc)
    max_bar = 3e-2 if n < 5000 else 1e-1
    assert rel_l2(y, ym[:, :n_out]) < 1e-3 and rel_max(y, ym[:, :n_out]) < 3e-3
    assert rel_l2(x1.grad, dxm[:, :n_in]) < 2e-3 and rel_max(x1.grad, dxm[:, :n_in]) < max_bar, "dx vs the stated arithmetic"
    assert rel_l2(net.params.grad, dWm) < 2e-3 and rel_max(net.params.grad, dWm) < max_bar, "dW vs the stated arithmetic"
    for a, b, name in [(x1.grad, x2.grad, 'dx'), (net.params.grad, p2.grad, 'dW')]:
        assert (a - b).abs().max().item() < 1e-3, name
        if n > 1000:
            assert rel_l2(a, b) < 5e-2, f"{name}: relative L2 error vs fp32 {rel_l2(a, b):.2e}"
    if n > 1000:
        record(f"mlp_wide_{n_in}_{hidden}x{n_hidden}_{n_out}", y_abs=e_y, dx_rel_l2_vs_fp32=rel_l2(x1.grad, x2.grad),
               dW_rel_l2_vs_fp32=rel_l2(net.params.grad, p2.grad))


def test_wide_mlp_weight_gradients_accumulate_across_splits():
    """Enough rows that a CTA of the split-K weight-gradient GEMM owns several consecutive splits of one output tile
    (8 tiles x 25 splits > 148 CTAs): those accumulate in TMEM and are flushed once.  Checked against the stated
    fp16 arithmetic and the fp32 oracle."""
    from autolabel_b200 import tcnn
    from oracle import field_oracle as fo
    from tests.helpers import rel_l2
    n, n_in, n_out, hidden, n_hidden = 50000 + 13, 15, 512, 512, 2
    net = tcnn.Network(n_in, n_out, {"otype": "CutlassMLP", "activation": "ReLU", "output_activation": "None",
                                     "n_neurons": hidden, "n_hidden_layers": n_hidden}).cuda()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(n, n_in, generator=g).cuda()
    x1 = x.clone().requires_grad_(True)
    y = net(x1)
    gy = torch.randn(n, n_out, generator=g).cuda() * 1e-4
    y.backward(gy)
    sc = fo.grad_scale_for(gy.abs().max().item())
    ym, dxm, dWm = fo.mlp_fp16_model(x, net.params.detach(), net.in_pad, hidden, net.out_pad, n_hidden, dout=gy, scale=sc)
    assert rel_l2(y, ym[:, :n_out]) < 1e-3
    assert rel_l2(x1.grad, dxm[:, :n_in]) < 2e-3
    assert rel_l2(net.params.grad, dWm) < 2e-3, rel_l2(net.params.grad, dWm)
    p2 = net.params.detach().clone().requires_grad_(True)
    oy = fo.mlp(x.clone(), p2, net.in_pad, hidden, net.out_pad, n_hidden)[:, :n_out]
    oy.backward(gy)
    assert rel_l2(net.params.grad, p2.grad) < 5e-2
