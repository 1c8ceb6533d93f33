"""result.h5 layout of the reference's training / evaluation logs (reference log.py:152-275).

What the reference writes: one resizable dataset per logged key under the file root -- shape (n_steps, *value_shape), grown by
one row per step -- holding only the keys that contain a whitelisted phrase (default ``'local_energy'``), static metadata as
file attributes, file opened in append mode with ``libver='v110'`` and SWMR switched on so that a reader can follow a running
job; on restart every dataset is cut back to ``init_step`` rows.  ``ResultTable`` / ``ResultLogger`` below produce exactly
that layout through the small slice of the h5py API they need (``File(path, 'a', libver=...)``, ``attrs.create``,
``create_dataset(name, shape, maxshape=, dtype=)``, ``Dataset.resize(n, axis=0)``, ``ds[-1, ...] = row``, ``visititems``).
h5py itself is NOT a dependency of this repository (it is absent from the build image): the logger imports it on first use
and says so if it is missing; the tests drive the same code through an in-memory stand-in with the same call surface
(tests/test_result_log.py), so the layout logic is checked without the library.  Values may be torch tensors (moved to the
host), numpy arrays or Python scalars.
"""
import os

import numpy as np


def flatten_stats(stats, parent='', sep='/'):
    """Nested statistics dict -> flat {'a/b/c': value} (the reference flattens with '/' before filtering, utils.py:216-224)."""
    flat = {}
    for key, value in stats.items():
        name = f'{parent}{sep}{key}' if parent else str(key)
        if isinstance(value, dict):
            flat.update(flatten_stats(value, name, sep))
        else:
            flat[name] = value
    return flat


def _to_host(value):
    if hasattr(value, 'detach'):  # torch tensor (any device)
        value = value.detach().cpu().numpy()
    if isinstance(value, (bool, int, float, np.generic)):
        return np.asarray(value)
    if isinstance(value, np.ndarray):
        return value
    raise ValueError(f'cannot log a value of type {type(value).__name__}')


class ResultTable:
    """Append-only table over an HDF5 group: ``table.append(key, value)`` adds one row to dataset ``key`` (created on first
    use as (0, *shape) with an unlimited first axis), ``table.truncate(n)`` cuts every dataset back to n rows, ``table[key]``
    reads a dataset (empty list if it does not exist yet) -- the reference's H5LogTable (log.py:164-201)."""

    def __init__(self, group):
        self.group = group

    def __getitem__(self, key):
        return self.group[key] if key in self.group else []

    def append(self, key, value):
        row = _to_host(value)
        if key not in self.group:
            # python floats are stored as float64, integers / arrays keep their dtype (as the reference's Appender does)
            self.group.create_dataset(key, (0, *row.shape), maxshape=(None, *row.shape), dtype=row.dtype)
        ds = self.group[key]
        if tuple(ds.shape[1:]) != tuple(row.shape):
            raise ValueError(f'{key}: row shape {row.shape} does not match the dataset rows {tuple(ds.shape[1:])}')
        ds.resize(ds.shape[0] + 1, axis=0)
        ds[-1, ...] = row

    def truncate(self, n_rows):
        def visit(name, obj):
            if hasattr(obj, 'resize') and hasattr(obj, 'shape'):  # datasets, not groups
                obj.resize(n_rows, axis=0)

        self.group.visititems(visit)


class ResultLogger:
    """``result.h5`` in ``workdir`` (reference H5Logger, log.py:204-275): ``update(stats)`` appends one row per whitelisted
    key of the (possibly nested) statistics of a step and flushes; ``init_step`` > 0 resumes: existing datasets are cut back
    to that many rows first.  ``aux_data`` becomes file attributes.  ``h5`` = the h5py module or a stand-in with its
    interface (default: import h5py)."""

    FILE_NAME = 'result.h5'
    DEFAULT_WHITELIST = ('local_energy',)

    def __init__(self, workdir, init_step=0, additional_keys_to_whitelist=None, aux_data=None, *, keys_whitelist=None, h5=None):
        if h5 is None:
            try:
                import h5py as h5
            except ImportError as exc:  # stated plainly: the layout code is here, the file-format library is not
                raise ImportError('ResultLogger needs h5py to write result.h5; it is not installed in this environment '
                                  '(pass h5=<module with the h5py interface> to use another backend)') from exc
        self.keys_whitelist = list(keys_whitelist if keys_whitelist is not None else self.DEFAULT_WHITELIST)
        self.keys_whitelist += list(additional_keys_to_whitelist or [])
        self.file = h5.File(os.path.join(workdir, self.FILE_NAME), 'a', libver='v110')
        self.file.swmr_mode = True  # single writer, readers may follow the running job
        for key, value in (aux_data or {}).items():
            self.file.attrs.create(key, value)
        self.table = ResultTable(self.file)
        self.table.truncate(init_step)
        self.flush()

    def update(self, stats):
        for key, value in flatten_stats(stats).items():
            if any(phrase in key for phrase in self.keys_whitelist):
                self.table.append(key, value)
        self.flush()

    def flush(self):
        self.file.flush()

    def close(self):
        self.file.close()
