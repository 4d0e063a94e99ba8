"""pytest plugin for the sub-process that runs the reference's own tests over the drop-in: at session end it reports
whether the native library was mapped into the process and how many fused launch chains ran, so that the outer test
can assert the B200 path (not some eager fallback) served those tests."""
import os


def pytest_sessionfinish(session, exitstatus):
    maps = open("/proc/self/maps").read() if os.path.exists("/proc/self/maps") else ""
    loaded = "libusflow_b200.so" in maps
    stats = "n/a"
    try:
        import ctypes
        from nf4ad_b200 import _lib
        st = (ctypes.c_longlong * 4)()
        buf = ctypes.create_string_buffer(160)
        _lib.lib().usf_debug_graph_stats(st, buf, 160)
        stats = f"graph_replays={st[0]} graph_captures={st[1]} capture_failures={st[2]} eager_runs={st[3]}"
    except Exception as e:          # report, never fail the inner session for this
        stats = f"unavailable ({type(e).__name__})"
    print(f"\nUSF_NATIVE_REPORT loaded={int(loaded)} {stats}")
