"""Host-side logic of bench.py that needs no GPU: the two-node pipeline harness."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_two_stage_pipeline_orders_and_gates():
    import bench
    log = []
    started = {}

    def stage_a(k, b):
        time.sleep(0.002)
        log.append(("a", k, b))
        return k * 10

    def stage_b(k, b, item):
        assert item == k * 10
        time.sleep(0.003)
        log.append(("b", k, b))

    def on_start():
        started["a_done"] = [x[1] for x in log if x[0] == "a"]
        started["b_done"] = [x[1] for x in log if x[0] == "b"]

    bench.run_two_stage_pipeline(12, 4, 3, stage_a, stage_b, on_start, lambda: log.append(("end",)), timeout_s=10.0)
    assert started["a_done"] == [0, 1, 2, 3] and started["b_done"] == [0, 1, 2, 3]      # warm-up fully drained before the clock starts
    assert [x[1] for x in log if x[0] == "b"] == list(range(12)) and log[-1] == ("end",)
    # never more than nbuf scans in flight
    inflight = 0
    for x in log:
        if x[0] == "a": inflight += 1
        elif x[0] == "b": inflight -= 1
        assert inflight <= 3
    bench.run_two_stage_pipeline(5, 0, 2, stage_a, stage_b, lambda: None, lambda: None, timeout_s=10.0)   # no warm-up


def test_two_stage_pipeline_propagates_errors():
    import bench
    import pytest

    def boom(k, b, item):
        raise ValueError("stage b failed")
    with pytest.raises(ValueError):
        bench.run_two_stage_pipeline(6, 2, 2, lambda k, b: k, boom, lambda: None, lambda: None, timeout_s=5.0)


def test_tools_and_entry_points_compile():
    """Every script a GPU session runs (tools/, bench.py, __graft_entry__.py, fixture generators) must at least parse."""
    import glob
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "tools", "*.py")) + [os.path.join(root, f) for f in ("bench.py", "__graft_entry__.py")] + \
        glob.glob(os.path.join(root, "tests", "make_golden*.py"))
    assert len(files) >= 8
    for f in files:
        py_compile.compile(f, doraise=True)
