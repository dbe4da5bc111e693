"""Queue helper (mirror of ``rlpyt/utils/synchronize.py:39-76``)."""
import queue


def drain_queue(queue_obj, n_sentinel=0, guard_sentinel=False):
    """Empty a multiprocessing queue; with ``n_sentinel`` block until that many ``None`` arrive."""
    contents = []
    if n_sentinel > 0:
        seen = 0
        while seen < n_sentinel:
            obj = queue_obj.get()
            if obj is None:
                seen += 1
            else:
                contents.append(obj)
        return contents
    while True:
        try:
            obj = queue_obj.get(block=False)
        except queue.Empty:
            return contents
        if guard_sentinel and obj is None:
            queue_obj.put(None)
            return contents
        if obj is not None:
            contents.append(obj)


class SpinSemaphore:
    """Single-producer / single-consumer counting semaphore in fork-shared memory that is waited on by
    SPINNING instead of sleeping on a futex (``multiprocessing.Semaphore``): a release is one store,
    an acquire sees it within a cache-line transfer, where a futex wake-up costs tens of microseconds
    of scheduler latency - twice per environment step per worker in the sampler's step loop
    (rlpyt/samplers/parallel/gpu/action_server.py:37-62 pairs every step with 2 x n_worker semaphore
    operations).  Same ``acquire() / acquire(block=False) / release()`` surface as the semaphores it
    replaces; exactly one process may release and exactly one may acquire a given instance.

    Correctness on x86-64 (TSO): ``posted`` is written only by the releaser and ``taken`` only by the
    acquirer, so no atomic read-modify-write is needed; stores become visible in program order, and the
    non-temporal observation stores of ``rl_host_stream_copy`` are followed by ``sfence`` before the
    counter is bumped.  A waiter that has spun for ``nap_after`` seconds starts napping between polls
    (a peer that is gone must not burn a core forever)."""

    def __init__(self, ctx=None, nap_after=2.0):
        import multiprocessing as mp
        import platform
        if platform.machine().lower() not in ("x86_64", "amd64"):
            raise RuntimeError("SpinSemaphore relies on x86-64 total store ordering (plain loads/stores, no fences); "
                               "use RLPYT_B200_SAMPLER_SYNC=futex on " + platform.machine())
        ctx = ctx or mp
        # two counters on separate cache lines (8 x int64 = 64 B apart)
        self._raw = ctx.RawArray("q", 16)
        self._nap_after = nap_after

    def _view(self):
        import numpy as np
        v = getattr(self, "_np", None)
        if v is None:
            v = self._np = np.frombuffer(self._raw, dtype=np.int64)
        return v

    def __getstate__(self):          # the numpy view is rebuilt lazily in the child / after pickling
        return {"_raw": self._raw, "_nap_after": self._nap_after}

    def __setstate__(self, state):
        self.__dict__.update(state)

    def release(self):
        v = self._view()
        v[0] += 1                    # posted

    def acquire(self, block=True, timeout=None):
        import time
        v = self._view()
        taken = v[8]
        if v[0] > taken:
            v[8] = taken + 1
            return True
        if not block:
            return False
        t0 = time.perf_counter()
        deadline = None if timeout is None else t0 + timeout
        spins = 0
        while v[0] <= taken:
            spins += 1
            if spins & 0xFFF == 0:
                now = time.perf_counter()
                if deadline is not None and now > deadline:
                    return False
                if now - t0 > self._nap_after:
                    time.sleep(0.0005)
        v[8] = taken + 1
        return True


class SpinThenSleepSemaphore:
    """A real ``multiprocessing.Semaphore`` with a bounded spin in front of it.

    ``release`` bumps a shared counter and posts the semaphore; ``acquire`` first polls the counter for
    at most ``spin_us`` microseconds and then ALWAYS takes the semaphore token - immediately if it was
    posted while spinning (no system call: glibc's sem_wait fast path), by sleeping on the futex
    otherwise.  Token accounting is exactly that of the plain semaphore (one post, one wait per
    handshake), so the protocol's blocking behaviour and its end-of-batch ``acquire(block=False)``
    drain checks are unchanged; only the futex sleep / wake-up latency of waits shorter than the spin
    bound disappears.  Single releaser and single acquirer per instance (like ``SpinSemaphore``)."""

    def __init__(self, ctx=None, spin_us=400.0):
        import multiprocessing as mp
        ctx = ctx or mp
        self._sem = ctx.Semaphore(0)
        self._raw = ctx.RawArray("q", 16)          # [0] posted, [8] taken: separate cache lines
        self._spin = spin_us * 1e-6

    def _view(self):
        import numpy as np
        v = self.__dict__.get("_np")
        if v is None:
            v = self.__dict__["_np"] = np.frombuffer(self._raw, dtype=np.int64)
        return v

    def release(self):
        v = self._view()
        v[0] += 1
        self._sem.release()

    def acquire(self, block=True, timeout=None):
        import time
        v = self._view()
        if not block:
            ok = self._sem.acquire(False)
            if ok:
                v[8] += 1
            return ok
        taken = v[8]
        if v[0] <= taken and self._spin > 0:
            t_end = time.perf_counter() + self._spin
            while v[0] <= taken and time.perf_counter() < t_end:
                pass
        ok = self._sem.acquire(True, timeout)
        if ok:
            v[8] = taken + 1
        return ok


class RWLock:
    """Multiple simultaneous readers, one writer - the ``RWLock`` of ``rlpyt/utils/synchronize.py:5-36`` (same
    surface: ``with lock:`` = read side, ``with lock.write_lock:`` = write side, ``acquire_read`` / ``release_read``
    / ``acquire_write`` / ``release_write``) for the asynchronous runner of this package, where sampler, copier and
    optimizer are THREADS of the one process that owns the GPU (the replay buffer and the parameters live in HBM;
    there is nothing to put into OS shared memory), so the primitives are ``threading`` ones."""

    def __init__(self):
        import threading
        self.write_lock = threading.Lock()
        self._read_lock = threading.Lock()
        self._read_count = 0

    def __enter__(self):
        self.acquire_read()

    def __exit__(self, *args):
        self.release_read()

    def acquire_write(self):
        self.write_lock.acquire()

    def release_write(self):
        self.write_lock.release()

    def acquire_read(self):
        with self._read_lock:
            self._read_count += 1
            if self._read_count == 1:
                self.write_lock.acquire()

    def release_read(self):
        with self._read_lock:
            self._read_count -= 1
            if self._read_count == 0:
                self.write_lock.release()


class StreamFence:
    """Orders device work issued by different host threads on different CUDA streams around one shared HBM object
    (the replay ring, the parameter staging copy).  The host-side lock alone is not enough on a GPU: kernels are only
    ENQUEUED while the lock is held.  ``after_write(stream)`` / ``after_read(stream)`` record an event on the issuing
    stream; ``before_read(stream)`` makes the stream wait for the last write, ``before_write(stream)`` for the last
    write and every read since.  Call them while holding the matching side of the ``RWLock``.  On a machine without
    CUDA (host-logic tests) every method is a no-op."""

    def __init__(self, device=None):
        import threading
        self._device = device
        self._mutex = threading.Lock()
        self._write_event = None
        self._read_events = {}

    @staticmethod
    def _stream(stream, device):
        import torch
        if not torch.cuda.is_available():
            return None
        return stream if stream is not None else torch.cuda.current_stream(device)

    def before_read(self, stream=None):
        s = self._stream(stream, self._device)
        if s is not None and self._write_event is not None:
            s.wait_event(self._write_event)

    def after_read(self, stream=None):
        s = self._stream(stream, self._device)
        if s is None:
            return
        ev = s.record_event()
        with self._mutex:
            self._read_events[s.cuda_stream] = ev              # a later event on the same stream covers the earlier ones

    def before_write(self, stream=None):
        s = self._stream(stream, self._device)
        if s is None:
            return
        if self._write_event is not None:
            s.wait_event(self._write_event)
        with self._mutex:
            reads, self._read_events = list(self._read_events.values()), {}
        for ev in reads:
            s.wait_event(ev)

    def after_write(self, stream=None):
        s = self._stream(stream, self._device)
        if s is not None:
            self._write_event = s.record_event()
