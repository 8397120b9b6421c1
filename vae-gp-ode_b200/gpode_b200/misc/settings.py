"""Global numeric settings (reference: experiments/model/misc/settings.py:5-34): fp32 everywhere,
parameters live on cuda:0 when a GPU is visible, jitter 1e-5."""
import numpy
import torch


class _Settings:
    torch_float = torch.float32
    numpy_float = numpy.float32
    torch_int = torch.int32
    numpy_int = numpy.int32
    jitter = 1e-5

    @property
    def device(self):
        return torch.device("cuda:0" if torch.cuda.is_available() else "cpu")


settings = _Settings()
