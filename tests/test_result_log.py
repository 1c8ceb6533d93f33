"""CPU: result.h5 layout logic (deepqmc_b200/log.py; reference log.py:152-275) through an in-memory stand-in with the slice of
the h5py interface the logger uses -- h5py is not installed here, and the logger must say so instead of failing obscurely."""
import numpy as np
import pytest
import torch

from deepqmc_b200 import log as L


class _Dataset:
    def __init__(self, shape, maxshape, dtype):
        assert maxshape[0] is None and tuple(maxshape[1:]) == tuple(shape[1:])
        self.data = np.zeros(shape, dtype=dtype if dtype is not None else np.float32)

    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)

    def resize(self, n, axis=0):
        assert axis == 0
        new = np.zeros((n, *self.data.shape[1:]), dtype=self.data.dtype)
        m = min(n, self.data.shape[0])
        new[:m] = self.data[:m]
        self.data = new

    def __setitem__(self, idx, value):
        self.data[idx] = value

    def __getitem__(self, idx):
        return self.data[idx]


class _Attrs(dict):
    def create(self, key, value):
        self[key] = value


class _File(dict):
    FILES = {}

    def __new__(cls, path, mode, libver=None):
        assert mode == 'a' and libver == 'v110'
        if path not in cls.FILES:  # append mode: re-opening sees what was written before
            cls.FILES[path] = super().__new__(cls)
            cls.FILES[path].attrs, cls.FILES[path].flushed, cls.FILES[path].swmr_mode = _Attrs(), 0, False
        return cls.FILES[path]

    def __init__(self, *a, **k):
        pass

    def create_dataset(self, name, shape, maxshape=None, dtype=None):
        self[name] = _Dataset(shape, maxshape, dtype)
        return self[name]

    def visititems(self, fn):
        for name, obj in self.items():
            fn(name, obj)

    def flush(self):
        self.flushed += 1

    def close(self):
        pass


class _H5:
    File = _File


def test_rows_are_appended_per_whitelisted_key_and_resume_truncates(tmp_path):
    lg = L.ResultLogger(str(tmp_path), 0, additional_keys_to_whitelist=['energy/mean'],
                        aux_data={'coords': np.eye(3)}, h5=_H5)
    f = lg.file
    assert f.swmr_mode is True and np.array_equal(f.attrs['coords'], np.eye(3))
    for step in range(5):
        lg.update({'local_energy': {'mol0': torch.full((2, 7), float(step))},      # nested + torch (any device)
                   'energy': {'mean': float(step), 'var': 1.0},                    # 'energy/var' is not whitelisted
                   'sampling': {'tau': np.float32(0.1)}})
    assert set(f.keys()) == {'local_energy/mol0', 'energy/mean'}
    assert f['local_energy/mol0'].shape == (5, 2, 7) and f['local_energy/mol0'].dtype == np.float32
    assert f['energy/mean'].shape == (5,) and f['energy/mean'].dtype == np.float64
    assert np.array_equal(f['energy/mean'][:], np.arange(5.0)) and f['local_energy/mol0'][3].min() == 3.0
    assert f.flushed == 6  # once at construction, once per step
    # restart from the checkpoint of step 3: every dataset is cut back, then grows again
    lg2 = L.ResultLogger(str(tmp_path), 3, additional_keys_to_whitelist=['energy/mean'], h5=_H5)
    assert lg2.file['energy/mean'].shape == (3,) and lg2.file['local_energy/mol0'].shape == (3, 2, 7)
    lg2.update({'local_energy': {'mol0': torch.zeros(2, 7)}, 'energy': {'mean': 9.0}})
    assert np.array_equal(lg2.file['energy/mean'][:], [0.0, 1.0, 2.0, 9.0])
    assert lg2.table['nothing yet'] == []
    with pytest.raises(ValueError):
        lg2.update({'local_energy': {'mol0': torch.zeros(3, 7)}})  # a row of another shape is an error, not a silent reshape
    with pytest.raises(ValueError):
        lg2.table.append('local_energy/bad', 'text')


def test_keys_whitelist_override_and_flatten():
    assert L.flatten_stats({'a': {'b': {'c': 1}, 'd': 2}, 'e': 3}) == {'a/b/c': 1, 'a/d': 2, 'e': 3}


def test_missing_h5py_is_reported_plainly(tmp_path):
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match='h5py'):
            L.ResultLogger(str(tmp_path))
    else:  # where the library exists the real file must have the reference's layout
        lg = L.ResultLogger(str(tmp_path), 0)
        lg.update({'local_energy': np.ones((4,), dtype=np.float32)})
        lg.close()
        with h5py.File(str(tmp_path / 'result.h5'), 'r') as f:
            assert f['local_energy'].shape == (1, 4) and f['local_energy'].maxshape == (None, 4)
