import datetime as dt


def unix2datetime(t):
    """seconds since epoch -> naive UTC datetime (what the reference's callers pass to Estimate)."""
    return dt.datetime.utcfromtimestamp(0) + dt.timedelta(seconds=float(t))
