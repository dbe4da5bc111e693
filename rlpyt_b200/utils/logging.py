"""A minimal tabular logger with the call surface the runners use of ``rlpyt/utils/logging/logger.py``
(``log``, ``set_iteration``, ``prefix`` / ``tabular_prefix`` contexts, ``record_tabular``, ``record_tabular_misc_stat``,
``dump_tabular``, ``save_itr_params``).  The reference's logger MODULE can be passed to the runners instead (it is
duck-typed); this one only prints and keeps the last table for tests.  Logging is outside the accelerated path
 - this exists so that the asynchronous runner can run without the reference installed."""
import contextlib
import time

import numpy as np


class TabularLogger:

    def __init__(self, stream=None, quiet=False):
        self._stream, self._quiet = stream, quiet
        self._prefixes, self._tab_prefixes = [], []
        self._rows = []
        self.iteration = 0
        self.last_table = {}
        self.tables = []
        self.snapshots = []

    def _emit(self, text):
        if not self._quiet:
            print(text, file=self._stream, flush=True)

    def log(self, s, *args, **kwargs):
        self._emit(time.strftime("%Y-%m-%d %H:%M:%S") + " | " + "".join(self._prefixes) + str(s))

    def set_iteration(self, iteration):
        self.iteration = iteration

    @contextlib.contextmanager
    def prefix(self, key):
        self._prefixes.append(key)
        try:
            yield
        finally:
            self._prefixes.pop()

    @contextlib.contextmanager
    def tabular_prefix(self, key):
        self._tab_prefixes.append(key)
        try:
            yield
        finally:
            self._tab_prefixes.pop()

    def record_tabular(self, key, val, *args, **kwargs):
        self._rows.append(("".join(self._tab_prefixes) + str(key), val))

    def record_tabular_misc_stat(self, key, values, placement="back"):
        """logger.py:452-470: Average / Std / Median / Min / Max of a list (NaN when empty)."""
        values = np.asarray(values, dtype=np.float64).reshape(-1) if len(values) > 0 else np.zeros(0)
        stats = (("Average", np.average), ("Std", np.std), ("Median", np.median), ("Min", np.min), ("Max", np.max))
        for name, fn in stats:
            label = f"{name}{key}" if placement == "front" else f"{key}{name}"
            self.record_tabular(label, float(fn(values)) if values.size else float("nan"))

    def dump_tabular(self, *args, **kwargs):
        table = dict(self._rows)
        self._rows = []
        self.last_table = table
        self.tables.append(table)
        if table:
            w = max(len(k) for k in table)
            self._emit("\n".join(f"{k.ljust(w)}  {v}" for k, v in table.items()))

    def save_itr_params(self, itr, params):
        self.snapshots.append(itr)          # snapshots are the caller's business; remember that one was requested
