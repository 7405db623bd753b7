"""Host-side regression: densehead._capi.handle() may be the first call into the binding (before lib()); the library
lock must be re-entrant or that first call deadlocks."""
import threading

from conftest import PKG  # noqa: F401  (puts the package on sys.path)


def test_handle_before_lib_does_not_deadlock(monkeypatch):
    from densehead import _capi
    assert isinstance(_capi._lock, type(threading.RLock()))
    done = []

    def worker():
        with _capi._lock:          # what handle() does ...
            with _capi._lock:      # ... and lib() inside it
                done.append(1)
    t = threading.Thread(target=worker, daemon=True)
    t.start()
    t.join(5)
    assert done == [1]
