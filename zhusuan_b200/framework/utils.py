"""Context stack (zhusuan/framework/utils.py:20-46) and ``reuse_variables``."""
import threading
from functools import wraps

__all__ = ["Context", "reuse_variables", "reuse"]


class Context(object):
    """Scoped "who is building right now" frames (the role of zhusuan/framework/utils.py:20-46).

    ``with frame:`` makes ``frame`` the innermost one of its class until the block ends;
    ``Cls.get_context()`` returns the innermost frame of ``Cls`` or raises ``RuntimeError``.
    Unlike the reference's class-level lists the stacks are PER THREAD (one dict of stacks in a
    ``threading.local``), so two host threads can build nets concurrently."""

    _tls = threading.local()

    @classmethod
    def get_contexts(cls):
        stacks = Context._tls.__dict__.setdefault("stacks", {})
        return stacks.setdefault(cls, [])       # Local and BayesianNet frames never mix

    @classmethod
    def get_context(cls):
        frames = cls.get_contexts()
        if not frames:
            raise RuntimeError("No contexts on the stack.")
        return frames[-1]

    def __enter__(self):
        type(self).get_contexts().append(self)
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        frames = type(self).get_contexts()
        assert frames and frames[-1] is self, "context frames must nest"
        frames.pop()


def reuse_variables(scope):
    """zhusuan/framework/utils.py:88-106 wraps ``tf.make_template`` so that TF
    variables created inside the function are shared between calls.  Device
    buffers here are explicit torch tensors / modules owned by the caller, so
    sharing is already the default; the decorator only keeps the call surface."""
    def wrapper(f):
        @wraps(f)
        def _wrapped(*args, **kwargs):
            return f(*args, **kwargs)
        _wrapped.scope = scope
        return _wrapped
    return wrapper


def reuse(scope):
    """(Deprecated) alias of :func:`reuse_variables` (framework/utils.py:109-117)."""
    import warnings
    warnings.warn(
        "The `reuse()` function has been renamed to `reuse_variables()`, "
        "`reuse()` will be removed in the coming version (0.4.1)",
        FutureWarning)
    return reuse_variables(scope)
