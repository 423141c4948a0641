"""Parameter holders of the candidate operations.

Same class names, constructor signatures, child-module names and creation order
as generalNAS_tools/operations_14_9.py:17-300, so that `state_dict()` keys,
`named_parameters()` order and seeded initialisation match the reference
tensor for tensor.  They hold weights and BatchNorm buffers only: the
arithmetic of every candidate is evaluated by the fused MixedOp / cell kernels
(csrc/cnn_fwd.cuh, csrc/cnn_bwd.cuh), never by these modules.
"""
import torch.nn as nn

PRIMITIVES_cnn = ['none', 'max_pool_5', 'avg_pool_5', 'skip_connect', 'sep_conv_15', 'sep_conv_9',
                  'dil_conv_15', 'dil_conv_9', 'conv_15']          # genotypes.py:29-39
PRIMITIVES_rnn = ['none', 'tanh', 'identity', 'sigmoid', 'relu']   # genotypes.py:42-48
rnn_steps = 8
CONCAT = 8


class _Holder(nn.Module):
    def forward(self, *args, **kwargs):
        raise RuntimeError(
            f"{type(self).__name__} only holds parameters: candidates are evaluated by the fused sm_100a "
            "MixedOp kernels (call the enclosing MixedOp / CNN_Cell_search / RNNModelSearch)")


def _conv(ci, co, k, stride=1, pad=0, dil=1, depthwise=False):
    return nn.Conv1d(ci, co, k, stride=stride, padding=pad, dilation=dil, groups=ci if depthwise else 1, bias=False)


def _bn(c, affine):
    return nn.BatchNorm1d(c, affine=affine)


class NorConv(_Holder):
    """conv_15: ReLU-conv(k,stride)-BN-ReLU-conv(k,1)-BN, children op.0..op.5 (operations_14_9.py:67-81)."""

    def __init__(self, C_in, C_out, kernel_size, stride, affine=True):
        super().__init__()
        self.op = nn.Sequential(nn.ReLU(), _conv(C_in, C_in, kernel_size, stride, 7), _bn(C_in, affine),
                                nn.ReLU(), _conv(C_in, C_out, kernel_size, 1, 7), _bn(C_out, affine))


class ReLUConvBN(_Holder):
    """preprocess: ReLU-conv-BN, children op.0..op.2 (operations_14_9.py:124-135)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, affine=True):
        super().__init__()
        self.op = nn.Sequential(nn.ReLU(), _conv(C_in, C_out, kernel_size, stride, padding), _bn(C_out, affine))


class DilConv(_Holder):
    """dil_conv_k: ReLU-depthwise(k,dilation)-pointwise-BN, children op.0..op.3 (operations_14_9.py:164-176)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, dilation, affine=True):
        super().__init__()
        self.op = nn.Sequential(nn.ReLU(), _conv(C_in, C_in, kernel_size, stride, padding, dilation, depthwise=True),
                                _conv(C_in, C_out, 1), _bn(C_out, affine))


class SepConv(_Holder):
    """sep_conv_k: two ReLU-depthwise-pointwise-BN blocks, children op.0..op.7 (operations_14_9.py:190-206)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding_1, padding_2, affine=True):
        super().__init__()
        layers = []
        for s, pad, co in ((stride, padding_1, C_in), (1, padding_2, C_out)):
            layers += [nn.ReLU(), _conv(C_in, C_in, kernel_size, s, pad, depthwise=True), _conv(C_in, co, 1), _bn(co, affine)]
        self.op = nn.Sequential(*layers)


class Identity(_Holder):
    """stride-1 skip_connect (operations_14_9.py:216)."""


class Zero(_Holder):
    """'none' (operations_14_9.py:242): contributes exactly zero to the mix and to every gradient."""

    def __init__(self, stride):
        super().__init__()
        self.stride = stride


class FactorizedReduce(_Holder):
    """strided skip_connect / preprocess0 after a reduction: ReLU, two strided 1x1 convs on the SAME
    positions, channel concat, BN (operations_14_9.py:280-300).  conv_1 and conv_2 are created back to
    back so that in the flat parameter buffer they form one [C_out, C_in] matrix."""

    def __init__(self, C_in, C_out, stride, affine=True):
        super().__init__()
        if C_out % 2:
            raise ValueError("FactorizedReduce needs an even number of output channels")
        self.relu = nn.ReLU()
        self.conv_1 = _conv(C_in, C_out // 2, 1, stride)
        self.conv_2 = _conv(C_in, C_out // 2, 1, stride)
        self.bn = _bn(C_out, affine)


def make_op(primitive, C, stride, affine=False):
    """Factory with the semantics of the reference's OPS table (operations_14_9.py:17-30)."""
    kind, _, k = primitive.rpartition('_')
    if primitive == 'none':
        return Zero(stride)
    if primitive == 'skip_connect':
        return Identity() if stride == 1 else FactorizedReduce(C, C, stride=stride, affine=affine)
    if kind == 'max_pool':
        return nn.MaxPool1d(int(k), stride=stride, padding=int(k) // 2)
    if kind == 'avg_pool':
        return nn.AvgPool1d(int(k), stride=stride, padding=int(k) // 2, count_include_pad=False)
    if kind == 'sep_conv':
        return SepConv(C, C, int(k), stride, (int(k) - 1) // 2, (int(k) - 1) // 2, affine=affine)
    if kind == 'dil_conv':
        return DilConv(C, C, int(k), stride, int(k) - 1, 2, affine=affine)
    if kind == 'conv':
        return NorConv(C, C, int(k), stride, affine=affine)
    raise KeyError(primitive)


OPS = {name: (lambda C, stride, affine, _n=name: make_op(_n, C, stride, affine)) for name in PRIMITIVES_cnn}
