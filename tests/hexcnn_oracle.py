"""The hex CNN of BASELINE config 5 (tools/hexcnn.py) evaluated with the CPU oracle's operators
(oracle/hexframes_oracle.py: the closed forms of HexConv2d / HexPool2d pinned against the live reference) and torch's own batch
norm / linear, on the product model's parameters.  TEST INFRASTRUCTURE: imported by tests/test_gpu_baseline_sizes.py only."""
import torch


class _RoundGradBf16(torch.autograd.Function):
    """Identity whose backward rounds the gradient to bfloat16: the tensor-core data / weight gradient kernels read the
    float32 gradient of a conv output through a bfloat16 conversion on its way into shared memory."""

    @staticmethod
    def forward(ctx, y):
        return y.view_as(y)

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _round_bf16(t):
    """Value rounded to bfloat16, gradient passed through unchanged (what a kernel that rounds its operand on load does)."""
    return t + (t.bfloat16().float() - t).detach()


def oracle_forward(params, x, eps=1e-5, autocast=False):
    """Forward of the same network on CPU tensors with the oracle's operators.  ``params``: dict name -> CPU tensor
    (requires_grad as wanted) with the product model's ``named_parameters()`` names; training-mode batch norm (batch
    statistics), conv without bias (HexConvModule's bias='auto' with a norm layer).  ``autocast``: restate where the
    product rounds to bfloat16 under torch.autocast -- every conv runs on the tensor cores (the RGB layer with its input
    channels rounded up to 16 in the loader) and reads bfloat16 activations, weights and output gradients, which is also
    what the reference's cuDNN convolution does under autocast."""
    from oracle import hexframes_oracle as HO
    F = torch.nn.functional
    for blk in ("c1", "c2", "c3"):
        w = params[f"{blk}.conv.kernel"]
        if autocast:
            x = _round_bf16(x)
            w = _round_bf16(w)
        x = HO.hexconv2d(x, w, None, 0, 2, 1, 1)
        if autocast:
            x = _RoundGradBf16.apply(x)
        x = F.batch_norm(x, None, None, params[f"{blk}.bn.weight"], params[f"{blk}.bn.bias"], True, 0.1, eps)
        x = F.relu(x)
        if blk != "c3":
            x = HO.hexpool2d(x, "max", 2, 2)
    x = HO.hexglobalpool2d(x, "average")
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):    # nn.Linear is an autocast op in the product too
        return F.linear(x, params["fc.weight"], params["fc.bias"])
