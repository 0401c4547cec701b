"""Global seed / draw counter for the in-kernel Philox generator (the
stand-in for ``tf.set_random_seed`` + TF's stateful op counters)."""
_state = {"seed": 0x5EED5EED, "counter": 0}


def set_random_seed(seed):
    _state["seed"] = int(seed) & 0xFFFFFFFFFFFFFFFF
    _state["counter"] = 0


def get_seed():
    return _state["seed"]


def next_counter():
    _state["counter"] = (_state["counter"] + 1) & 0xFFFFFFFF
    return _state["counter"]


# ---- CUDA-graph replay of a sampling step -----------------------------------------------------
# A step captured once would freeze the (seed, counter) pairs its samplers were called with; a
# registered device epoch is added to the counter INSIDE the kernels and bumped on the stream at
# the end of every (replayed) step, so each replay draws fresh numbers (include/zsb200.h:
# zsb_random_set_device_epoch).
_epoch = {"tensor": None}


def enable_device_epoch(device="cuda"):
    """Allocate and register the device epoch (idempotent); returns the uint32 tensor."""
    import torch
    from ._lib import lib, ptr
    if _epoch["tensor"] is None:
        t = torch.zeros(1, dtype=torch.int32, device=device)
        lib.call("zsb_random_set_device_epoch", ptr(t))
        _epoch["tensor"] = t
    return _epoch["tensor"]


def disable_device_epoch():
    from ._lib import lib
    if _epoch["tensor"] is not None:
        lib.call("zsb_random_set_device_epoch", None)
        _epoch["tensor"] = None


def bump_device_epoch(by):
    """Advance the device epoch by ``by`` on the current stream (capturable)."""
    from ._lib import lib, ptr, stream
    lib.call("zsb_random_bump_epoch", ptr(_epoch["tensor"]), int(by), stream())


def counter():
    return _state["counter"]


def set_counter(value):
    _state["counter"] = int(value) & 0xFFFFFFFF
