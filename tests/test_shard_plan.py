"""CPU: the shard plan the library uses (ms_shard_plan -> csrc/comm.cuh shard_range / subtree_plan) against the Python
planners of ministark_b200/sharded.py that the gloo tests exercise."""
import ctypes as C

import pytest


def test_shard_plan_matches_the_python_planner():
    from ministark_b200 import _lib
    from ministark_b200.sharded import SubtreePlan, column_ranges

    lib = _lib.load()
    for world in (1, 2, 3, 4, 8, 16):
        for cols in (1, 2, 3, 5, 8, 16, 31, 32, 64):
            want = column_ranges(cols, world)
            for rank in range(world):
                a, b = C.c_uint64(), C.c_uint64()
                lib.ms_shard_plan(cols, 1, 2, world, rank, C.byref(a), C.byref(b), None, None)
                assert (a.value, b.value) == want[rank]
        for k in (2, 4, 8, 16):
            for lg in range(0, 14):
                groups = 1 << lg
                per, left = C.c_uint64(), C.c_uint64()
                rc = lib.ms_shard_plan(1, groups, k, world, 0, None, None, C.byref(per), C.byref(left))
                try:
                    plan = SubtreePlan.make(groups, k, world)
                except ValueError:
                    assert rc != 0, (groups, k, world)
                    continue
                assert rc == 0 and (per.value, left.value) == (plan.groups_per_rank, plan.digests_per_rank), (groups, k, world)


def test_nccl_error_code_is_reachable():
    """MS_ERR_NCCL (include/ministark.h): a host whose NCCL cannot be loaded gets error 3 from the communicator entry
    points instead of a crash (fresh process: the library caches the dlopen)."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import ctypes as C; from ministark_b200 import _lib; lib = _lib.load(); b = (C.c_uint8 * 128)(); "
            "print(lib.ms_comm_unique_id(b))")
    env = dict(os.environ, MINISTARK_NCCL_LIB="/nonexistent/libnccl.so.2", PYTHONPATH=root)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, cwd=root)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "3"
