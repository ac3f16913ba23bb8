"""bench.py's Lifeline: the JSON line survives a rank that fails, or hangs, inside an optional multi-rank object."""
import json
import os
import subprocess
import sys
import time

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import sys, time
sys.path.insert(0, %(repo)r)
import bench
rank, key, mode = int(sys.argv[1]), sys.argv[2], sys.argv[3]
line = {"metric": "overlap_reads_per_s", "value": 1.0, "cfg4": {"index_ms": 1.0}}
life = bench.Lifeline(rank, line, key, poll_s=0.05)
time.sleep(0.5)  # stands for the barrier behind the constructor
life.arm("cfg5", 1.5 if mode == "deadline" else 60.0)
try:
    if mode == "fail" and rank == 1:
        raise RuntimeError("boom")
    if mode == "ok":
        life.disarm()
        if rank == 0:
            print("{}")
        sys.exit(0)
    time.sleep(60)  # a collective the failed rank never joins
except Exception as ex:
    life.fail("cfg5", "%%s: %%s" %% (type(ex).__name__, ex))
print("not reached")
"""


@pytest.mark.parametrize("mode", ["fail", "deadline", "ok"])
def test_lifeline_prints_the_line_once_and_every_rank_exits_zero(tmp_path, mode):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"repo": REPO})
    key = "test_%d_%s" % (os.getpid(), mode)
    t0 = time.time()
    procs = [subprocess.Popen([sys.executable, str(script), str(r), key, mode], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(3)]
    outs = [p.communicate(timeout=40) for p in procs]
    assert time.time() - t0 < 30
    assert [p.returncode for p in procs] == [0, 0, 0], [o[1][-500:] for o in outs]
    assert outs[1][0] == "" and outs[2][0] == ""
    lines = outs[0][0].splitlines()
    assert len(lines) == 1
    line = json.loads(lines[0])
    flag = os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", "siga_b200_bench_abort_" + key)
    if mode == "ok":
        assert line == {} and not os.path.exists(flag)
        return
    assert line["value"] == 1.0 and line["cfg4"] == {"index_ms": 1.0}  # what was complete is kept
    if mode == "fail":
        assert "rank 1" in line["cfg5"]["error"] and "boom" in line["cfg5"]["error"]
        os.remove(flag)
    else:
        assert "deadline" in line["cfg5"]["error"]
