"""Regression guard for the stuck replay of round 2 (profiles/r02_replay_stall_probe.txt): 2000 back-to-back replays of the
graphed perception step, with the L2 flush between them like bench.py, must all complete.  The loop runs in a CHILD process
(tools/hang_probe.py) so that a stuck kernel dies with its context instead of taking the test session's device with it."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_thousand_replays_complete():
    cmd = [sys.executable, os.path.join(ROOT, "tools", "hang_probe.py"), "--replays", "2000", "--seed", "8", "--flush",
           "--tag", "pytest"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    except subprocess.TimeoutExpired:
        pytest.fail("tools/hang_probe.py did not return within 240 s")
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not lines:
        pytest.skip(f"probe produced no verdict (rc {r.returncode}): {r.stderr[-400:]}")
    v = json.loads(lines[-1])
    assert not v["hung"], f"the device stopped making progress at replay {v['at_replay']} of {v['replays']}"
    assert v["at_replay"] == 2000
