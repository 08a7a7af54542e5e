"""Import-time stand-in for `humanfriendly` (absent offline); only what the reference touches at import."""


def format_timespan(seconds, *args, **kwargs):
    return '{:.2f} seconds'.format(float(seconds))


def format_size(n, *args, **kwargs):
    return '{} bytes'.format(n)
