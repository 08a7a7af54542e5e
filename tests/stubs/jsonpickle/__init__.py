"""Minimal stand-in for the `jsonpickle` package (absent offline; the reference imports it at module load time in
utils/ct_utils.py and uses it only for pretty-printing objects).  Test infrastructure."""
import json


def set_encoder_options(*args, **kwargs):
    pass


def encode(obj, *args, **kwargs):
    return json.dumps(obj, default=lambda o: getattr(o, '__dict__', str(o)))


def decode(s, *args, **kwargs):
    return json.loads(s)
