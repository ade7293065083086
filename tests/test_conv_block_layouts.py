"""The conv-bn-lrelu block of the host modules has two execution paths over ONE set of parameters
(planar_optical_flow_b200/model/dr_spaam.py::_ConvBnAct): the reference's NCL path (dr_spaam.py:8-12) and the
channels-last 4-D path the training branch uses, which leaves the convolution bias out of the forward pass under batch
statistics.  They must be the same function: outputs, running statistics, gradients.  CPU only."""
import pytest
import torch

from planar_optical_flow_b200.model.dr_spaam import _conv


def _run(block, x, four_d):
    if four_d:
        x4 = x.unsqueeze(2).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        y = block(x4)
        y.square().mean().backward()
        return y.squeeze(2), x4.grad.squeeze(2)
    x3 = x.clone().requires_grad_(True)
    y = block(x3)
    y.square().mean().backward()
    return y, x3.grad


@pytest.mark.parametrize("cin,cout,k,pad,L", [(1, 8, 3, 1, 12), (8, 16, 3, 1, 7), (16, 4, 5, 0, 5)])
@pytest.mark.parametrize("training", [True, False])
def test_channels_last_path_is_the_same_function(cin, cout, k, pad, L, training):
    torch.manual_seed(cin * 100 + cout)
    a, b = _conv(cin, cout, k, pad), _conv(cin, cout, k, pad)
    with torch.no_grad():
        a[0].bias.normal_()                          # a non-trivial bias: it must reach the running mean
        a[1].weight.uniform_(0.5, 1.5)
        a[1].bias.normal_()
        a[1].running_mean.normal_()
        a[1].running_var.uniform_(0.5, 2.0)
    b.load_state_dict(a.state_dict())
    a.train(training)
    b.train(training)
    x = torch.randn(6, cin, L)
    for _ in range(2):                               # two steps: the running statistics feed the second
        ya, ga = _run(a, x, four_d=False)
        yb, gb = _run(b, x, four_d=True)
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6)
    assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-6)
    for (na, pa), (nb, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert na == nb
        if na == "0.bias" and training:              # zero in exact arithmetic on both paths (rounding noise on the NCL path)
            assert float(pa.grad.abs().max()) < 1e-5 and float(pb.grad.abs().max()) == 0.0
        else:
            assert torch.allclose(pa.grad, pb.grad, rtol=1e-4, atol=1e-6), na
    for (na, ba), (nb, bb) in zip(a.named_buffers(), b.named_buffers()):
        assert na == nb and torch.allclose(ba.float(), bb.float(), rtol=1e-5, atol=1e-6), na
    assert list(a.state_dict().keys()) == ["0.weight", "0.bias", "1.weight", "1.bias", "1.running_mean", "1.running_var",
                                           "1.num_batches_tracked"]
