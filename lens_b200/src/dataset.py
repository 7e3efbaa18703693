"""Query dataset with the reference's interface (lens/src/dataset.py).

`CustomImageDataset(annotations_file, img_dir, kernel_size, transform, ..., skip, max_samples,
is_spiking, time_window)` lists the same files in the same order (CSV column 0, `iloc[::skip]`,
`iloc[:max_samples]`, dataset.py:76-84).  The reference's `__getitem__` builds a
[T, 1, roi, roi] float Bernoulli raster on the CPU (dataset.py:118-125); here the raster is
generated inside the CUDA feature kernel from the raw u8 frame, so the fast path only needs
`load_frames()`.  `__getitem__` is kept for callers that iterate the dataset the reference's
way: it returns the same 4-tuple, with the raster computed by torch (host data preparation,
exactly the reference's own expressions).
"""
import math
import os

import numpy as np
import pandas as pd
import torch
from torch.utils.data import Dataset


def read_png_u8(path):
    """Greyscale PNG -> u8 [H, W] (torchvision.io.read_image(...)[0] in the reference)."""
    try:
        from torchvision.io import read_image
        img = read_image(path)
        return img[0].contiguous()
    except ImportError:  # pragma: no cover - torchvision is present in the target image
        from PIL import Image
        return torch.from_numpy(np.array(Image.open(path).convert("L")))


class ProcessImage:
    """lens/src/dataset.py:28-51 with is_train=False: u8 image -> flat float pixel / 255."""

    def __init__(self, is_train=False):
        if is_train:
            raise NotImplementedError("training transforms are out of scope")
        self.intensity = 255

    def __call__(self, img):
        n = img.shape[0]
        flat = img.view(n, 1, -1) / self.intensity
        return torch.squeeze(torch.squeeze(flat, 0), 0)


class CustomImageDataset(Dataset):
    def __init__(self, annotations_file, img_dir, kernel_size, transform=None, target_transform=None,
                 skip=1, max_samples=None, test=True, is_spiking=False, time_window=33):
        self.transform, self.target_transform = transform, target_transform
        self.skip, self.time_window, self.is_spiking = skip, time_window, is_spiking
        self.kernel_size, self.test = kernel_size, test
        labels = pd.read_csv(annotations_file)
        labels["file_path"] = [os.path.join(img_dir, str(n)) for n in labels.iloc[:, 0]]
        labels = labels.iloc[::skip]
        if max_samples is not None:
            labels = labels.iloc[:max_samples]
        self.img_labels = labels

    def __len__(self):
        return len(self.img_labels)

    def paths(self):
        return list(self.img_labels["file_path"])

    def labels(self):
        return list(self.img_labels.iloc[:, 1])

    def load_frames(self):
        """All query frames as one u8 tensor [Q, H, W] (host)."""
        frames = []
        for p in self.paths():
            if not os.path.exists(p):
                raise FileNotFoundError(f"No file found at {p}.")
            frames.append(read_png_u8(p))
        return torch.stack(frames)

    def __getitem__(self, idx):
        row = self.img_labels.iloc[idx]
        path = row["file_path"]
        if not os.path.exists(path):
            raise FileNotFoundError(f"No file found for index {idx} at {path}.")
        image = read_png_u8(path)[None]
        label = self.img_labels.iloc[idx, 1]
        third = self.img_labels.iloc[idx, 2]
        if self.transform:
            image = self.transform(image)
        if self.target_transform:
            label = self.target_transform(label)
        if self.is_spiking:
            torch.manual_seed(50)
            image = (torch.rand(self.time_window, *image.shape) < image).float()
            side = int(math.sqrt(image[-1].size()[0]))
            image = image.view(self.time_window, side, side).unsqueeze(1)
        return image, label, third, []
