"""Replay buffers for the asynchronous runner (mirror of ``rlpyt/replays/async_.py:8-47`` ``AsyncReplayBufferMixin``
and of the ``Async*`` buffer classes of ``rlpyt/replays/non_sequence/frame.py`` / ``replays/sequence/frame.py``).

The reference keeps the ring in OS shared memory, guards it with a multi-process read-write lock and publishes the
cursor through shared ``RawValue``s, because its writers (memory-copier processes) and its reader (the optimizer
process) are different processes.  Here the ring is in HBM and writer and reader are threads of the process that owns
the GPU, so the cursor is simply the object's attribute; what has to be added instead is DEVICE ordering: ``append`` is
enqueued on the copier's stream and ``sample_batch`` / ``update_batch_priorities`` on the optimizer's, and holding the
host lock while ENQUEUEING does not order the kernels.  ``StreamFence`` makes each stream wait for the other's last
conflicting operation (events, no host synchronisation)."""
from rlpyt_b200.utils.synchronize import RWLock, StreamFence


class AsyncReplayBufferMixin:

    async_ = True

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.rw_lock = RWLock()
        self._fence = StreamFence(getattr(self, "device", None))

    def append_samples(self, *args, **kwargs):
        with self.rw_lock.write_lock:
            self._fence.before_write()
            ret = super().append_samples(*args, **kwargs)
            self._fence.after_write()
        return ret

    def sample_batch(self, *args, **kwargs):
        with self.rw_lock:                                   # read lock
            self._fence.before_read()
            batch = super().sample_batch(*args, **kwargs)
            self._fence.after_read()
        return batch

    def update_batch_priorities(self, *args, **kwargs):
        with self.rw_lock.write_lock:
            self._fence.before_write()
            ret = super().update_batch_priorities(*args, **kwargs)
            self._fence.after_write()
        return ret
