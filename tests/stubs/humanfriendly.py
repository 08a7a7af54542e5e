"""Minimal stand-in for the `humanfriendly` package (absent offline; the reference imports it at module load time for
log messages only: run_detector.py, run_detector_batch.py).  Test infrastructure."""


def format_timespan(seconds, *args, **kwargs):
    return '{:.2f} seconds'.format(float(seconds))


def format_size(n, *args, **kwargs):
    return '{} bytes'.format(int(n))
