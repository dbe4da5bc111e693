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
