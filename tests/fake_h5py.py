"""A stand-in for the few h5py features the snapshot writer / reader use (File as a context manager,
groups, attrs, datasets), backed by pickle files: h5py is not in this image, and the write -> read round trip
is host logic worth testing without it.  Test infrastructure only; the product imports the real h5py."""
import os
import pickle

import numpy as np


class Dataset:
    def __init__(self, data):
        self.data = np.array(data)

    def __getitem__(self, idx):
        return self.data[idx]

    def __array__(self, dtype=None, copy=None):
        return self.data if dtype is None else self.data.astype(dtype)

    @property
    def shape(self):
        return self.data.shape


class Group:
    def __init__(self):
        self.attrs = {}
        self.items_ = {}

    def create_group(self, name):
        if name in self.items_:
            raise ValueError(f"group {name} exists")
        g = Group()
        self.items_[name] = g
        return g

    def create_dataset(self, name, data=None):
        d = Dataset(data)
        self.items_[name] = d
        return d

    def __getitem__(self, name):
        return self.items_[name]          # KeyError like h5py

    def __contains__(self, name):
        return name in self.items_

    def __iter__(self):
        return iter(sorted(self.items_))  # h5py iterates in name order

    def __len__(self):
        return len(self.items_)


class File(Group):
    def __init__(self, filename, mode="r"):
        super().__init__()
        self.filename, self.mode = filename, mode
        if mode == "r":
            if not os.path.exists(filename):
                raise FileNotFoundError(filename)
            with open(filename, "rb") as fh:
                root = pickle.load(fh)
            self.attrs, self.items_ = root.attrs, root.items_

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.mode == "w":
            root = Group()
            root.attrs, root.items_ = self.attrs, self.items_
            with open(self.filename, "wb") as fh:
                pickle.dump(root, fh)
        return False
