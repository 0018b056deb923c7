"""Drop-in for the reference's utils/pytorch_correlation.py: same class name and
constructor, routed to the fused sm_100a correlation kernel (the reference's
F.unfold formulation, utils/pytorch_correlation.py:27-50, survives only as the
oracle in oracle/ref_port.py)."""
from upflow_pytorch_b200 import ops
from utils.tools import tools


class Corr_pyTorch(tools.abstract_model):
    def __init__(self, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1):
        # same argument checks as the reference (utils/pytorch_correlation.py:17-18)
        assert pad_size == max_displacement
        assert stride1 == stride2 == 1
        super().__init__()
        if kernel_size != 1:
            raise NotImplementedError("only kernel_size=1 is used by UPFlow (model/upflow.py:354-355)")
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.stride1 = stride1
        self.stride2 = stride2
        self.max_hdisp = max_displacement

    def forward(self, in1, in2):
        return ops.correlation(in1, in2, self.max_hdisp)
