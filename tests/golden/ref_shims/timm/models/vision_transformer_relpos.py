import torch.nn as nn
class RelPosAttention(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("stub")
