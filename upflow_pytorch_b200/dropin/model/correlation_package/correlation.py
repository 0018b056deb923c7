"""Drop-in for model/correlation_package/correlation.py (the Python face of the
reference's `correlation_cuda` extension, correlation.py:6-61): `Correlation`
and `CorrelationFunction` with the same constructor arguments, backed by
libupflow_b200.so (upf_corr_lrelu_fwd / upf_corr_lrelu_bwd)."""
from torch.nn.modules.module import Module

from upflow_pytorch_b200 import ops


def _check(pad_size, kernel_size, max_displacement, stride1, stride2):
    if kernel_size != 1 or stride1 != 1 or stride2 != 1 or pad_size != max_displacement:
        raise NotImplementedError(
            "upflow_b200 correlation supports the configuration UPFlow uses (model/upflow.py:561): "
            "kernel_size=1, stride1=stride2=1, pad_size==max_displacement; got k=%s s1=%s s2=%s pad=%s maxd=%s"
            % (kernel_size, stride1, stride2, pad_size, max_displacement))


class CorrelationFunction:
    """Callable with the legacy (constructor-configured) interface of correlation.py:6-44."""

    def __init__(self, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2, corr_multiply=1):
        _check(pad_size, kernel_size, max_displacement, stride1, stride2)
        self.max_displacement = max_displacement

    def __call__(self, input1, input2):
        return ops.correlation(input1, input2, self.max_displacement)


class Correlation(Module):
    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super(Correlation, self).__init__()
        _check(pad_size, kernel_size, max_displacement, stride1, stride2)
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        return ops.correlation(input1, input2, self.max_displacement)
