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
