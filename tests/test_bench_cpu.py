"""Host-side pieces of bench.py that can be checked without a GPU: the clock sampler's parsing / time-window filter."""
import datetime
import importlib.util
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


class _Proc:
    def terminate(self):
        pass


class _Thread:
    def join(self, timeout=None):
        pass


def _row(now, dt, mhz, power_cap="Not Active"):
    ts = datetime.datetime.fromtimestamp(now + dt).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
    return [ts, str(mhz), "1965", "500.0", "Not Active", "Not Active", "Not Active", power_cap]


def test_clock_sampler_filters_to_the_timed_region():
    b = _bench()
    now = time.time()
    s = b.ClockSampler(0)
    s.proc, s.thread = _Proc(), _Thread()
    s.rows = [_row(now, -1.0, 1000), _row(now, 0.01, 1965), _row(now, 0.05, 1960, "Active"), _row(now, 0.3, 900)]
    out = s.stop(now, now + 0.1)
    assert out["samples"] == 2 and out["in_timed_region"] and out["sm_mhz"] == 1962.5
    assert out["reasons"] == ["sw_power_cap"] and out["sm_max_mhz"] == 1965.0


def test_clock_sampler_survives_garbage():
    b = _bench()
    now = time.time()
    s = b.ClockSampler(0)
    s.proc, s.thread = _Proc(), _Thread()
    s.rows = [["garbage"], ["x", "y"]]
    assert s.stop(now, now + 0.1) is None
    s.proc = _Proc()
    s.rows = [["not a timestamp", "1950", "1965", "1", "Not Active", "Not Active", "Not Active", "Not Active"]]
    out = s.stop(now, now + 0.1)
    assert out["samples"] == 1 and out["in_timed_region"] is False
