import torch.nn as nn


def convert_syncbn_model(m, *a, **k):
    return m


class DistributedDataParallel(nn.Module):
    def __init__(self, module, **k):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)
