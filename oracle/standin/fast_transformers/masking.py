import torch


class TriangularCausalMask:
    def __init__(self, N, device="cpu"):
        self.N = N
        self.lower_triangular = True


class LengthMask:
    def __init__(self, lengths, max_len=None, device=None):
        self.lengths = lengths
        self.max_len = max_len or int(lengths.max())

    @property
    def float_matrix(self):
        idx = torch.arange(self.max_len)[None, :]
        return (idx < self.lengths[:, None]).float()
