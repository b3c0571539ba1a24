"""In-memory stand-in for the subset of the h5py / zarr group API that the FlacArray file layout uses.

h5py and zarr are optional dependencies of the reference (hdf5_utils.py:13-21, zarr.py:13-22) and are
not installed in every environment.  `MemGroup` implements the same protocol -- `attrs`, `in`,
`__getitem__`, `create_dataset` (h5py spelling) and `create_array` (zarr spelling), datasets with
`shape / dtype / size / attrs`, numpy-style slicing and h5py's `read_direct / write_direct` -- so that
the layout code in `hdf5.py` / `zarr.py` can be exercised (and used as an in-memory container) without
either library.  With the real libraries installed the same functions take real groups.
"""
import numpy as np


class MemDataset:
    def __init__(self, shape, dtype, zarr_style=False):
        self._data = np.zeros(tuple(int(x) for x in shape), dtype=np.dtype(dtype))
        self.attrs = dict()
        if not zarr_style:
            # h5py-only members; a zarr array has neither, and the I/O code must cope with that
            self.read_direct = self._read_direct
            self.write_direct = self._write_direct

    @property
    def shape(self):
        return self._data.shape

    @property
    def dtype(self):
        return self._data.dtype

    @property
    def size(self):
        return self._data.size

    def __getitem__(self, key):
        return np.array(self._data[key])

    def __setitem__(self, key, value):
        self._data[key] = value

    def _read_direct(self, dest, source_sel=None, dest_sel=None):
        src = self._data if source_sel is None else self._data[source_sel]
        if dest_sel is None:
            dest[...] = src
        else:
            dest[dest_sel] = src

    def _write_direct(self, source, source_sel=None, dest_sel=None):
        src = source if source_sel is None else source[source_sel]
        if dest_sel is None:
            self._data[...] = src
        else:
            self._data[dest_sel] = src


class MemGroup:
    """Dictionary-backed group.  `zarr_style=True` hides the h5py-only members of its datasets."""

    def __init__(self, zarr_style=False):
        self.attrs = dict()
        self._items = dict()
        self._zarr_style = zarr_style

    def __contains__(self, name):
        return name in self._items

    def __getitem__(self, name):
        return self._items[name]

    def keys(self):
        return self._items.keys()

    def create_dataset(self, name, shape=None, dtype=None, data=None, **kwargs):
        if data is not None:
            data = np.asarray(data)
            shape, dtype = data.shape, data.dtype
        ds = MemDataset(shape, dtype, zarr_style=self._zarr_style)
        if data is not None:
            ds[...] = data
        self._items[name] = ds
        return ds

    def create_array(self, name, shape=None, dtype=None, **kwargs):
        return self.create_dataset(name, shape=shape, dtype=dtype)

    def create_group(self, name):
        g = MemGroup(zarr_style=self._zarr_style)
        self._items[name] = g
        return g

    def require_group(self, name):
        return self._items[name] if name in self._items else self.create_group(name)
