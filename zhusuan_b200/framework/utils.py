"""Context stack (zhusuan/framework/utils.py:20-46) and ``reuse_variables``."""
from functools import wraps

__all__ = ["Context", "reuse_variables", "reuse"]


class Context(object):
    """Context stack; class-level lists, one per subclass, NOT thread-safe
    (same as the reference, framework/utils.py:35-46)."""
    _contexts = {}

    def __enter__(self):
        type(self).get_contexts().append(self)
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        type(self).get_contexts().pop()

    @classmethod
    def get_contexts(cls):
        # one stack per class (Local and BayesianNet do not share a stack)
        return Context._contexts.setdefault(cls, [])

    @classmethod
    def get_context(cls):
        try:
            return cls.get_contexts()[-1]
        except IndexError:
            raise RuntimeError("No contexts on the stack.")


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
