import torch.nn as nn
class LightningModule(nn.Module):
    pass
