"""Host allocator tuning for the test / bench processes.

In the sandboxes this repo runs in, first-touch page faults cost ~10-200 ms per MiB, and glibc hands every freed
large block straight back to the OS (mmap threshold), so numpy temporaries are re-faulted over and over.  Raising the
mmap / trim thresholds keeps freed blocks in the heap for reuse.  Input synthesis only — no effect on device code."""
import ctypes

_done = False


def tune_malloc() -> bool:
    global _done
    if _done:
        return True
    try:
        libc = ctypes.CDLL("libc.so.6")
        ok = libc.mallopt(-3, 1 << 30) and libc.mallopt(-1, 2 ** 31 - 1)  # M_MMAP_THRESHOLD, M_TRIM_THRESHOLD
        _done = bool(ok)
    except OSError:
        _done = False
    return _done
