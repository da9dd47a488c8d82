"""Dataset helpers with the reference's names (reference src/data_utils.py).  I/O is outside the built hot path
(SURVEY.md section 2 row 7): the benchmark uses synthetic tensors; these classes only have to exist, load the
reference's file formats and hand out [C, P, P] float tensors."""
from pathlib import Path

import numpy as np
import pandas as pd
import torch
from torch.utils import data

from .grad_utils import generalized_b_xy_c_to_image, generalized_image_to_b_xy_c  # noqa: F401


def cycle(dl):
    while True:
        for d in dl:
            yield d


class Dataset(data.Dataset):
    """CSV fields -> [N, C, P, P] held in RAM (reference data_utils.py:31-75): one CSV per channel, one flattened
    sample per row."""

    def __init__(self, data_directories, use_double=False, return_img=True, gaussian_prior=False):
        super().__init__()
        dtype = torch.float64 if use_double else torch.float32
        chans = []
        for path in data_directories:
            arr = pd.read_csv(path, header=None).to_numpy()
            p = int(np.sqrt(arr.shape[1]))
            chans.append(torch.tensor(arr, dtype=dtype).reshape(-1, 1, p, p))
        self.data = torch.cat(chans, dim=1)
        self.return_img = return_img

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        x = self.data[index]
        return x if self.return_img else generalized_image_to_b_xy_c(x[None])[0]


class Dataset_Paths(data.Dataset):
    """One .npy per sample, [10, 65, 65] (reference data_utils.py:77-119)."""

    def __init__(self, data_directory, use_double=False):
        super().__init__()
        self.paths = sorted(Path(data_directory).glob('*.npy'))
        self.dtype = torch.float64 if use_double else torch.float32

    def __len__(self):
        return len(self.paths)

    def __getitem__(self, index):
        return torch.tensor(np.load(self.paths[index]), dtype=self.dtype)
