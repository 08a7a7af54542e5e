"""Import-time stand-in for `jsonpickle` (absent offline); only what the reference touches at import."""
import json


def set_encoder_options(*args, **kwargs):
    pass


def encode(obj, *args, **kwargs):
    return json.dumps(obj, default=lambda o: getattr(o, '__dict__', str(o)))


def decode(s, *args, **kwargs):
    return json.loads(s)
